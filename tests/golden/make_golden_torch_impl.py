"""Generate golden vectors by IMPORTING the reference's own pure-PyTorch twin.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_torch_impl.py

Writes small .npz fixtures next to this file.  They pin the CPU oracle (oracle/gstex_oracle.c)
to the reference's ``gstex_cuda/_torch_impl.py``:

* ``torch_impl_sample.npz``  - ``_torch_impl.sample_texture`` on the inputs of the reference's
  tests/test_sample.py:9-34 (seeded here; upstream is unseeded).
* ``torch_impl_sh.npz``      - ``_torch_impl.compute_sh_color`` and its autograd coefficient
  gradient for degrees 0..4 (reference tests/test_sh.py:9-46 pins degree 4).
* ``torch_impl_raster_*.npz``- ``_torch_impl.texture_forward`` plus torch-autograd gradients of the
  example.py:189-209 loss, on example.py's initialisation (``example.py:69-119``) at the
  ``--torch_compare`` sizes.  The reference's torch path calls CUDA binning
  (texture.py:470), which cannot run here, so the tile lists come from the CPU oracle's
  binning restatement; the inputs and the lists are stored in the fixture.

The reference code is imported from where it lies; nothing is copied.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import gstex_cuda._torch_impl as _T  # noqa: E402  (the reference, imported in place)

import oracle  # noqa: E402


def gen_sample():
    g = torch.Generator().manual_seed(7)
    sz = 5
    dims = torch.stack([
        torch.randint(6, (sz,), generator=g, dtype=torch.int32) + 2,
        torch.randint(7, (sz,), generator=g, dtype=torch.int32) + 2,
        torch.zeros((sz,), dtype=torch.int32)], dim=-1)
    hws = dims[:, 0] * dims[:, 1]
    dims[:, -1] = torch.cumsum(hws, 0) - hws
    total = int(hws.sum())
    ch = 10
    texture = torch.rand((total, ch), generator=g)
    nq = 100
    uvs = torch.rand((nq, 2), generator=g)
    # include exact borders and out-of-range queries (clamped by the reference)
    uvs[:6] = torch.tensor([[0., 0.], [1., 1.], [1., 0.3], [0.2, 1.], [-.25, .5], [.5, 1.75]])
    ids = torch.randint(sz, (nq,), generator=g)
    qdims = dims[ids]
    out = _T.sample_texture(qdims, texture, uvs)
    np.savez_compressed(os.path.join(HERE, "torch_impl_sample.npz"), texture_dims=qdims.numpy(),
                        texture=texture.numpy(), uvs=uvs.numpy(), out=out.numpy())


def gen_sh():
    g = torch.Generator().manual_seed(11)
    n = 64
    d = {}
    dirs = torch.randn(n, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    d["viewdirs"] = dirs.numpy()
    for deg in range(5):
        K = (deg + 1) ** 2
        coeffs = torch.rand(n, K, 3, generator=g, requires_grad=True)
        colors = _T.compute_sh_color(dirs, coeffs)
        v = torch.randn(n, 3, generator=g)
        (colors * v).sum().backward()
        d[f"coeffs{deg}"] = coeffs.detach().numpy()
        d[f"colors{deg}"] = colors.detach().numpy()
        d[f"v_colors{deg}"] = v.numpy()
        d[f"v_coeffs{deg}"] = coeffs.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "torch_impl_sh.npz"), **d)


def example_scene(seed, H, W, num_points, th, tw, background):
    """example.py:69-119 initialisation + :121-143 preprocess, on CPU."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    N = num_points
    means = torch.rand(N, 3) - 0.5
    means[:, :2] *= 16
    raw_scales = 0.5 * np.log(1 / N) * torch.rand(N, 3)
    raw_rgbs = torch.rand(N, 3)
    u, v, w = torch.rand(N, 1), torch.rand(N, 1), torch.rand(N, 1)
    raw_quats = torch.cat([
        torch.sqrt(1.0 - u) * torch.sin(2.0 * math.pi * v), torch.sqrt(1.0 - u) * torch.cos(2.0 * math.pi * v),
        torch.sqrt(u) * torch.sin(2.0 * math.pi * w), torch.sqrt(u) * torch.cos(2.0 * math.pi * w)], -1)
    raw_opac = torch.ones((N, 1))
    mapping = torch.zeros((N, 1, 4))
    mapping[:, :, :2] = 0.5
    sc = 0.25 * np.sqrt(1 / torch.sum(torch.exp(raw_scales[:, 0] + raw_scales[:, 1])).item())
    mapping[:, :, 2] = np.log(sc)
    mapping[:, :, 3] = 2.0 * math.pi * torch.rand(N, 1)
    raw_texture = torch.rand(N * th * tw, 3)
    viewmat = torch.eye(4)
    viewmat[2, 3] = 8.0
    c2w = viewmat.inverse()
    focal = 0.5 * float(W) / math.tan(0.5 * math.pi / 2.0)

    scales = torch.zeros_like(raw_scales)
    scales[:, :2] = torch.exp(raw_scales[:, :2])
    scales[:, -1] = 1e-5 * torch.mean(scales[:, :-1], dim=-1)
    quats = raw_quats / raw_quats.norm(dim=-1, keepdim=True)
    Rs = _T.normalized_quat_to_rotmat(quats)
    uv0 = mapping[:, :, :2].clone()
    uvscale = torch.exp(mapping[:, :, None, 2])
    theta = mapping[:, :, None, 3]
    ax1, ax2 = Rs[:, None, :, 0], Rs[:, None, :, 1]
    umap = uvscale * (ax1 * torch.cos(theta) + ax2 * torch.sin(theta))
    vmap = uvscale * (-ax1 * torch.sin(theta) + ax2 * torch.cos(theta))
    dims = torch.zeros(N, 3, dtype=torch.int32)
    dims[:, 0], dims[:, 1] = th, tw
    dims[:, 2] = torch.cumsum(dims[:, 0] * dims[:, 1], 0) - dims[:, 0] * dims[:, 1]
    return dict(
        means=means, scales=scales, quats=quats, colors=torch.sigmoid(raw_rgbs), opacities=torch.sigmoid(raw_opac),
        uv0=uv0, umap=umap, vmap=vmap, texture=torch.sigmoid(raw_texture), texture_dims=dims, viewmat=viewmat,
        c2w=c2w, intrins=(focal, focal, W / 2, H / 2), background=torch.tensor(background, dtype=torch.float32),
        H=H, W=W)


def gen_raster(name, seed, H, W, N, th, tw, background, settings=1 << 8):
    s = example_scene(seed, H, W, N, th, tw, background)
    bw = 16
    npy = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in s.items()}
    b = oracle.bin_view(npy["means"], npy["scales"], 1.0, npy["quats"], npy["viewmat"], s["intrins"], H, W, bw)
    leaves = {}
    for k in ("colors", "opacities", "means", "scales", "quats", "uv0", "umap", "vmap", "texture"):
        leaves[k] = s[k].clone().requires_grad_(True)
    fx, fy, cx, cy = s["intrins"]
    outs = _T.texture_forward(
        b["tile_bounds"], (bw, bw, 1), (W, H, 1), (N, 1, 3), s["texture_dims"],
        torch.from_numpy(b["gaussian_ids_sorted"]), torch.from_numpy(b["tile_bins"]), leaves["colors"],
        leaves["opacities"], leaves["means"], leaves["scales"], 1.0, leaves["quats"], leaves["uv0"], leaves["umap"],
        leaves["vmap"], leaves["texture"], s["viewmat"], s["c2w"], fx, fy, cx, cy, settings, s["background"])
    out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, final_idx = outs
    out_alpha = 1 - final_Ts
    # example.py:189-209 loss against the default red/blue/white target, plus small terms that exercise
    # the out_img / alpha / depth gradient inputs (the example's loss leaves those at zero).
    gt = torch.ones((H, W, 3))
    gt[: H // 2, : W // 2, :] = torch.tensor([1.0, 0.0, 0.0])
    gt[H // 2:, W // 2:, :] = torch.tensor([0.0, 0.0, 1.0])
    loss = torch.nn.functional.mse_loss(out_texture, gt) + out_reg.mean() + (
        out_normal[:, :, 0] ** 2 + out_normal[:, :, 1] ** 2 + (1 - out_normal[:, :, 2]) ** 2).mean()
    loss = loss + 0.3 * torch.nn.functional.mse_loss(out_img, 1 - gt) + 0.2 * (out_alpha ** 2).mean() \
        + 0.01 * out_depth.mean()
    outs_req = [out_img, out_depth, out_reg, out_alpha, out_texture, out_normal]
    for o in outs_req:
        o.retain_grad()
    loss.backward()
    d = dict(
        H=H, W=W, block_width=bw, settings=settings, glob_scale=1.0, intrins=np.array(s["intrins"], np.float32),
        viewmat=npy["viewmat"], c2w=npy["c2w"], background=npy["background"], texture_dims=npy["texture_dims"],
        gaussian_ids_sorted=b["gaussian_ids_sorted"], tile_bins=b["tile_bins"], loss=float(loss))
    for k, v in leaves.items():
        d[k] = v.detach().numpy()
        d["v_" + k] = v.grad.numpy() if v.grad is not None else np.zeros_like(v.detach().numpy())
    for k, o in zip(("out_img", "out_depth", "out_reg", "out_alpha", "out_texture", "out_normal"), outs_req):
        d[k] = o.detach().numpy()
        d["v_" + k] = o.grad.numpy() if o.grad is not None else np.zeros_like(o.detach().numpy())
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "loss", float(loss), "M", b["num_intersects"])


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_sample()
    gen_sh()
    # BASELINE config 1 (example.py --height 32 --width 32 --num_points 10, seed 1) with 11x11 texels per
    # Gaussian instead of 317x317 so that the fixture stays small.
    gen_raster("torch_impl_raster_c1.npz", seed=1, H=32, W=32, N=10, th=11, tw=11, background=[0., 0., 0.])
    gen_raster("torch_impl_raster_b.npz", seed=5, H=48, W=48, N=40, th=5, tw=3, background=[0.2, 0.5, 0.9])
    gen_raster("torch_impl_raster_nouv.npz", seed=9, H=32, W=32, N=24, th=4, tw=4, background=[1., 1., 1.],
               settings=0)
