"""-m gpu: BASELINE.json configs 1-3 at their exact shapes (example.py's initialisation, camera and settings), CUDA path
against the CPU oracle, forward and backward:

  C1  example.py --height 32 --width 32 --num_points 10          -> 317x317 texels per Gaussian (1 M texels / 10)
  C2  example.py defaults: 256x256, 100 Gaussians                 -> 101x101 texels per Gaussian
  C3  example.py --num_points 10000 --num_texels 0                -> 1x1 texels (pure 2DGS with the texture path on)

(Config 4 at full size is tests/test_gpu_full_size.py; config 5 is tests/test_gpu_pipeline.py + test_data_parallel_gloo.py.)
"""
import pytest
import torch

from gpu_util import (DEV, to_np, bin_cuda, forward_cuda, backward_cuda, forward_oracle, backward_oracle, random_vout,
                      compare_forward, compare_backward)
from test_gpu_train_ops import example_raw, quat_axes12

pytestmark = pytest.mark.gpu


def example_scene(N, H, W, num_texels, seed):
    """example.py:69-143: raw leaves -> the activated tensors SimpleTrainer.forward hands to the rasteriser."""
    th = tw = int((num_texels / N) ** 0.5 + 1)  # example.py:91-92
    raw, dims, intr = example_raw(N, H, W, th, tw, seed)
    raw["quats"] = raw["quats"] / raw["quats"].norm(dim=-1, keepdim=True)
    raw["opacities"] = torch.ones_like(raw["opacities"])          # example.py:93
    raw["texture"] = torch.rand(raw["texture"].shape, generator=torch.Generator().manual_seed(seed + 1)).to(DEV)
    scales = torch.zeros_like(raw["scales"])
    scales[:, :2] = torch.exp(raw["scales"][:, :2])
    scales[:, 2] = 1e-5 * scales[:, :2].mean(dim=-1)
    a1, a2 = quat_axes12(raw["quats"])
    mp = raw["mapping"]
    us, th_ = torch.exp(mp[:, :, None, 2]), mp[:, :, None, 3]
    umap = (us * (a1[:, None] * torch.cos(th_) + a2[:, None] * torch.sin(th_))).squeeze(2)
    vmap = (us * (-a1[:, None] * torch.sin(th_) + a2[:, None] * torch.cos(th_))).squeeze(2)
    vm = torch.eye(4, device=DEV)
    vm[2, 3] = 8.0
    s = dict(means=raw["means"], scales=scales.contiguous(), quats=raw["quats"].contiguous(),
             opacities=torch.sigmoid(raw["opacities"]), colors=torch.sigmoid(raw["rgbs"]),
             texture=torch.sigmoid(raw["texture"]), texture_dims=dims, uv0=mp[:, :, :2].contiguous(),
             umap=umap.contiguous(), vmap=vmap.contiguous(), viewmat=vm, c2w=torch.linalg.inv(vm).contiguous(),
             background=torch.zeros(3, device=DEV), intrins=intr, H=H, W=W, num_points=N, texture_info=(N, 1, 3),
             glob_scale=1.0, settings=1 << 8, block_width=16)
    return s, (th, tw)


@pytest.mark.parametrize("name,N,H,W,num_texels,texdim", [("C1", 10, 32, 32, 1_000_000, 317),
                                                          ("C2", 100, 256, 256, 1_000_000, 101),
                                                          ("C3", 10000, 256, 256, 0, 1)])
def test_baseline_config_vs_oracle(name, N, H, W, num_texels, texdim):
    s, (th, tw) = example_scene(N, H, W, num_texels, seed=1)
    assert th == tw == texdim and s["texture"].shape[0] == N * texdim * texdim
    b = bin_cuda(s)
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    f_c, scratch = forward_cuda(s, ids, bins)
    f_o = forward_oracle(s, to_np(ids), to_np(bins))
    print(f"{name}: M = {b['num_intersects']}, covered = {float((f_o['final_Ts'] < 0.99).mean()):.2f}")
    assert float((f_o["final_Ts"] < 0.99).mean()) > 0.05
    compare_forward(f_c, f_o, max_bad_frac=2e-3, int_bad_frac=2e-3)
    vout = random_vout(s, 7)
    g_c = backward_cuda(s, ids, bins, f_c, vout, scratch=scratch)
    g_o = backward_oracle(s, to_np(ids), to_np(bins), f_c, vout)
    # example.py draws uniformly random orientations, so at 10 000 Gaussians a few surfels are seen edge-on (C3: Gaussian
    # 950 has cos(normal, ray) = 0.0015).  Their plane denominator is a difference of O(1) terms: an fp32 evaluation of
    # it - the reference kernel's dot(ray, ax3), the oracle's - carries ~1e-4 relative error that the 1/D^2 factors of the
    # quaternion / mean gradients amplify (the oracle differs from the reference CUDA kernel by 2.1 % of max|g| on such
    # entries).  Here the form coefficients of grazing surfels are computed in double (csrc/pack.cu), so what remains is
    # the oracle's own noise: the standard outlier bound applies to C3 like to every other case.
    uv_keys = ("v_uv0", "v_umap", "v_vmap")
    if texdim == 1:
        # 1x1 textures: all four bilinear corners are the same texel, so d(out)/d(uv) vanishes identically; both sides
        # return fp32 cancellation noise (~1e-6 against gradients of ~1e2) instead of a comparable signal
        for k in uv_keys:
            assert float(g_c[k].abs().max()) <= 1e-4 and float(abs(g_o[k]).max()) <= 1e-4
        g_c = {k: (torch.zeros_like(v) if k in uv_keys else v) for k, v in g_c.items()}
        g_o = {k: (0 * v if k in uv_keys else v) for k, v in g_o.items()}
    compare_backward(g_c, g_o, max_bad_frac=2e-3, outlier_bound=0.05)
