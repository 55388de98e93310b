"""Pure-PyTorch helper functions with the names of the reference's ``gstex_cuda/_torch_impl.py``.

These are the small differentiable helpers callers import next to the rasteriser (``example.py:11`` uses
``normalized_quat_to_rotmat``).  They run on whatever device their inputs live on and are NOT part of the
accelerated path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)
_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
       -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def eval_sh_bases(basis_dim: int, dirs: Tensor) -> Tensor:
    """_torch_impl.py:62-113: real SH basis (1, 4, 9, 16 or 25 functions) at UNIT directions."""
    x, y, z = dirs.unbind(-1)
    cols = [torch.full_like(x, _C0)]
    if basis_dim > 1:
        cols += [-_C1 * y, _C1 * z, -_C1 * x]
    if basis_dim > 4:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        cols += [_C2[0] * xy, _C2[1] * yz, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * xz, _C2[4] * (xx - yy)]
    if basis_dim > 9:
        cols += [_C3[0] * y * (3 * xx - yy), _C3[1] * xy * z, _C3[2] * y * (4 * zz - xx - yy),
                 _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
                 _C3[6] * x * (xx - 3 * yy)]
    if basis_dim > 16:
        cols += [_C4[0] * xy * (xx - yy), _C4[1] * yz * (3 * xx - yy), _C4[2] * xy * (7 * zz - 1),
                 _C4[3] * yz * (7 * zz - 3), _C4[4] * (zz * (35 * zz - 30) + 3), _C4[5] * xz * (7 * zz - 3),
                 _C4[6] * (xx - yy) * (7 * zz - 1), _C4[7] * xz * (xx - 3 * yy),
                 _C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(cols[:basis_dim], dim=-1)


def compute_sh_color(viewdirs: Tensor, sh_coeffs: Tensor) -> Tensor:
    """_torch_impl.py:12-22: (*, D, C) coefficients -> (*, C) colours (no +0.5, no clamp)."""
    bases = eval_sh_bases(sh_coeffs.shape[-2], viewdirs)
    return (bases[..., None] * sh_coeffs).sum(dim=-2)


def normalized_quat_to_rotmat(quat: Tensor) -> Tensor:
    """_torch_impl.py:116-133: rotation matrices of unit quaternions (w, x, y, z)."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(quat, dim=-1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def quat_to_rotmat(quat: Tensor) -> Tensor:
    return normalized_quat_to_rotmat(F.normalize(quat, dim=-1))


def project_pix(fxfy, p_view, center, eps=1e-6):
    """_torch_impl.py:140-147: pinhole projection with +1e-6 on z."""
    rw = 1.0 / (p_view[..., 2] + 1e-6)
    return torch.stack([p_view[..., 0] * rw * fxfy[0] + center[0], p_view[..., 1] * rw * fxfy[1] + center[1]], dim=-1)


def sample_texture(texture_dim: Tensor, texture: Tensor, uvs: Tensor) -> Tensor:
    """_torch_impl.py:149-194: bilinear fetch with replicate padding from the jagged texture."""
    h, w, start = texture_dim[:, 0], texture_dim[:, 1], texture_dim[:, 2]
    tu = h * uvs[:, 0].clamp(0, 1)
    tv = w * uvs[:, 1].clamp(0, 1)
    i0f, j0f = torch.floor(tu), torch.floor(tv)
    fu, fv = (tu - i0f)[:, None], (tv - j0f)[:, None]
    i0, j0 = i0f.to(torch.int64), j0f.to(torch.int64)
    h64, w64, s64 = h.to(torch.int64), w.to(torch.int64), start.to(torch.int64)
    i1, j1 = torch.minimum(i0 + 1, h64 - 1), torch.minimum(j0 + 1, w64 - 1)
    i0, j0 = torch.minimum(i0, h64 - 1), torch.minimum(j0, w64 - 1)
    t00, t01 = texture[s64 + i0 * w64 + j0], texture[s64 + i0 * w64 + j1]
    t10, t11 = texture[s64 + i1 * w64 + j0], texture[s64 + i1 * w64 + j1]
    return (1 - fu) * (1 - fv) * t00 + (1 - fu) * fv * t01 + fu * (1 - fv) * t10 + fu * fv * t11


def texture_forward(tile_bounds, block, img_size, texture_info, texture_dims, gaussian_ids_sorted, tile_bins, colors,
                    opacities, means3d, scales, glob_scale, quats, uv0s, umaps, vmaps, texture, viewmat, c2w, fx, fy, cx,
                    cy, settings, background):
    """Pure-PyTorch, autograd-differentiable rasteriser with the semantics of the reference's CPU/GPU twin
    (``_torch_impl.py:196-381``) - the slow checker behind ``texture_gaussians(..., use_torch_impl=True)`` and
    ``example.py --torch_compare``.  NOT the accelerated path.  Like the reference twin (and unlike the kernels) it caps
    alpha at 0.999, zeroes the alpha of skipped pairs BEFORE the stop test, has no blur floor and no NDC distortion, and
    reports ``final_idx`` = the end of the tile's list.  Returns the 7-tuple
    ``(out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, final_idx)``, all (H, W, ...).

    Organisation: rays of the whole frame once, then one dense (Gaussians x pixels) evaluation per tile."""
    W, H = int(img_size[0]), int(img_size[1])
    bw_x, bw_y = int(block[0]), int(block[1])
    dev, f32 = colors.device, torch.float32
    C_tex = int(texture_info[-1])
    out_img = (torch.ones((H, W, colors.shape[1]), dtype=f32, device=dev) * background).contiguous()
    out_depth, out_reg = torch.zeros((H, W), dtype=f32, device=dev), torch.zeros((H, W), dtype=f32, device=dev)
    out_texture, out_normal = torch.zeros((H, W, C_tex), dtype=f32, device=dev), torch.zeros((H, W, 3), dtype=f32, device=dev)
    final_Ts, final_idx = torch.ones((H, W), dtype=f32, device=dev), torch.zeros((H, W), dtype=torch.int32, device=dev)

    origin = c2w[:3, 3]
    ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    dirs_cam = torch.stack([(xs + 0.5 - cx) / fx, (ys + 0.5 - cy) / fy, torch.ones_like(xs, dtype=f32)], -1)
    rays_all = dirs_cam @ c2w[:3, :3].T
    rays_all = rays_all / rays_all.norm(dim=-1, keepdim=True)                       # (H, W, 3) unit world rays
    vdep_all = rays_all @ viewmat[2, :3]                                            # view depth per unit ray length
    Rs = normalized_quat_to_rotmat(quats)
    bins = tile_bins.tolist()
    propagate_uv = bool(settings & (1 << 8))

    for ty in range(int(tile_bounds[1])):
        for tx in range(int(tile_bounds[0])):
            lo, hi = bins[ty * int(tile_bounds[0]) + tx]
            if hi <= lo:
                continue
            y0, y1, x0, x1 = ty * bw_y, min(ty * bw_y + bw_y, H), tx * bw_x, min(tx * bw_x + bw_x, W)
            rays = rays_all[y0:y1, x0:x1].reshape(-1, 3)                            # (P, 3)
            vdep = vdep_all[y0:y1, x0:x1].reshape(-1)
            g = gaussian_ids_sorted[lo:hi].long()
            R, mean, sc = Rs[g], means3d[g], glob_scale * scales[g]
            a1, a2, a3 = R[:, :, 0], R[:, :, 1], R[:, :, 2]
            # ray-plane intersection with the +-1e-6 denominator clamp (texture_helpers.cuh:302-313)
            den = a3 @ rays.T                                                       # (G, P)
            den = torch.where((den >= 0) & (den < 1e-6), torch.full_like(den, 1e-6), den)
            den = torch.where((den <= 0) & (den > -1e-6), torch.full_like(den, -1e-6), den)
            t = ((a3 * (mean - origin)).sum(-1))[:, None] / den
            delta = origin[None, None, :] + t[:, :, None] * rays[None, :, :] - mean[:, None, :]
            l1, l2 = (delta * a1[:, None, :]).sum(-1), (delta * a2[:, None, :]).sum(-1)
            sigma = 0.5 * (l1 * l1 / sc[:, None, 0] ** 2 + l2 * l2 / sc[:, None, 1] ** 2)
            alpha = torch.clamp(opacities[g] * torch.exp(-sigma), max=0.999)
            alpha = torch.where((alpha < 1.0 / 255) | (t < 0.01) | (t > 1000.0), torch.zeros_like(alpha), alpha)
            # transmittance in front of each Gaussian; a Gaussian whose blend would leave T <= 1e-4 is not blended
            T_after = torch.cumprod(1 - alpha, dim=0)
            T_before = torch.cat([torch.ones_like(alpha[:1]), T_after[:-1]], 0)
            T = torch.where(T_after <= 1e-4, torch.zeros_like(T_before), T_before)
            vis = alpha * T
            acc = vis.sum(0)
            d_uv = delta if propagate_uv else delta.detach()
            uu = torch.clamp(uv0s[g][:, 0, None, 0] + (d_uv * umaps[g]).sum(-1), 0.0, 1.0)
            vv = torch.clamp(uv0s[g][:, 0, None, 1] + (d_uv * vmaps[g]).sum(-1), 0.0, 1.0)
            if not propagate_uv:
                uu, vv = uu.detach(), vv.detach()
            P = rays.shape[0]
            tdi = texture_dims[g][:, None, :].expand(-1, P, -1).reshape(-1, 3)
            samples = sample_texture(tdi, texture, torch.stack([uu, vv], -1).reshape(-1, 2)).reshape(len(g), P, -1)
            hh, ww = y1 - y0, x1 - x0
            out_img[y0:y1, x0:x1] = ((vis[:, :, None] * colors[g][:, None, :]).sum(0)
                                     + (1 - acc)[:, None] * background[None, :]).reshape(hh, ww, -1)
            out_normal[y0:y1, x0:x1] = (vis[:, :, None] * a3[:, None, :]).sum(0).reshape(hh, ww, 3)
            out_texture[y0:y1, x0:x1] = (vis[:, :, None] * samples).sum(0).reshape(hh, ww, -1)
            zero = torch.zeros_like(vis[:1])
            s0 = torch.cat([zero, vis[:-1]], 0).cumsum(0)
            s1 = torch.cat([zero, (vis * t)[:-1]], 0).cumsum(0)
            s2 = torch.cat([zero, (vis * t * t)[:-1]], 0).cumsum(0)
            out_reg[y0:y1, x0:x1] = (vis * (t * t * s0 + s2 - 2 * t * s1)).sum(0).reshape(hh, ww)
            # median depth: the last Gaussian (in list order) with alpha > 1e-3 in front of which T > 0.5
            seen = ((alpha > 1e-3) & (T > 0.5)).int() * (1 + torch.arange(len(g), device=dev))[:, None]
            val, pos = seen.max(0)
            depth = torch.where(val != 0, (t * vdep[None, :])[pos, torch.arange(P, device=dev)], torch.zeros_like(vdep))
            out_depth[y0:y1, x0:x1] = depth.reshape(hh, ww)
            final_Ts[y0:y1, x0:x1] = (1 - acc).reshape(hh, ww)
            final_idx[y0:y1, x0:x1] = hi
    return out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, final_idx
