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
