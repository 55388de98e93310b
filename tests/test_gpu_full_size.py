"""-m gpu: BASELINE config 4 at FULL size (1 M textured Gaussians, SH degree 3, 4x4 texels, 1920x1080) checked through
size-independent properties - the CPU oracle cannot finish this size in test time:

  binning   keys sorted, tile ranges partition [0, M), every listed Gaussian's AABB touches its tile, the fused
            bucket-by-tile binning equals the staged cumsum -> emit -> radix sort -> edges path bit for bit
  forward   transmittance in [0, 1], indices inside the tile's range, bit-identical when repeated, LINEAR in the colours
            and in the texture (same geometry): F(a c1 + b c2) = a F(c1) + b F(c2)
  backward  LINEAR in the upstream gradients, and the adjoint identity <J^T v, d> = <v, J d> for the two linear inputs
            (colours -> out_img, texture -> out_texture), which ties the backward kernel to the forward kernel at full size
"""
import numpy as np
import pytest
import torch

from gstex_cuda_b200 import sh as SH
from gstex_cuda_b200 import utils as U
from gstex_cuda_b200.scenes import synthetic_scene
from gpu_util import DEV, bin_cuda, forward_cuda, backward_cuda, random_vout

pytestmark = pytest.mark.gpu

N, W, H, BW = 1_000_000, 1920, 1080, 16


@pytest.fixture(scope="module")
def c4():
    s = synthetic_scene(N, W, H, seed=1234, device=DEV)
    s["colors"] = SH.spherical_harmonics_colors(3, s["means"], s["c2w"], s["sh_coeffs"]).contiguous()
    b = bin_cuda(s, BW)
    torch.cuda.synchronize()
    return s, b


def test_full_size_binning_properties(c4):
    s, b = c4
    m, tb = b["num_intersects"], b["tile_bounds"]
    assert 2_000_000 < m < 6_000_000  # the C4 scene: about 3.3 M intersections
    keys, ids, bins = b["isect_ids_sorted"], b["gaussian_ids_sorted"], b["tile_bins"]
    assert bool((keys[1:] >= keys[:-1]).all())
    sizes = (bins[:, 1] - bins[:, 0]).to(torch.int64)
    assert int(sizes.sum()) == m and int(sizes.min()) >= 0
    nz = bins[sizes > 0]
    assert int(nz[0, 0]) == 0 and int(nz[-1, 1]) == m and bool((nz[1:, 0] == nz[:-1, 1]).all())  # a partition of [0, M)
    tile_of = torch.repeat_interleave(torch.arange(bins.shape[0], device=DEV), sizes)
    assert bool(((keys >> 32) == tile_of).all())
    # every listed Gaussian's screen AABB really touches its tile
    tx, ty = (tile_of % tb[0]).float(), (tile_of // tb[0]).float()
    c, e = b["centers"][ids.long()], b["extents"][ids.long()]
    assert bool(((c[:, 0] + e[:, 0] >= tx * BW - 1e-3) & (c[:, 0] - e[:, 0] <= (tx + 1) * BW + 1e-3)
                 & (c[:, 1] + e[:, 1] >= ty * BW - 1e-3) & (c[:, 1] - e[:, 1] <= (ty + 1) * BW + 1e-3)).all())
    # fused bucket-by-tile binning == staged path, bit for bit
    ids_f, bins_f, count_f, keys_f = U.bin_tiles(b["centers"], b["extents"], b["depths"], tb, BW, m + 1000, want_isect_ids=True)
    assert int(count_f.item()) == m
    assert torch.equal(ids_f[:m], ids) and torch.equal(bins_f, bins) and torch.equal(keys_f[:m], keys)


def test_full_size_forward_properties_and_linearity(c4):
    s, b = c4
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    s0 = dict(s, background=torch.zeros(3, device=DEV))
    f, _ = forward_cuda(s0, ids, bins)
    f2, _ = forward_cuda(s0, ids, bins)
    for k in f:
        assert torch.equal(f[k], f2[k]), f"forward not reproducible: {k}"
        assert bool(torch.isfinite(f[k].float()).all()), k
    T = f["final_Ts"]
    assert float(T.min()) >= 0.0 and float(T.max()) <= 1.0
    assert float((1 - T).mean()) > 0.5  # the scene covers the frame
    # final_idx / depth_idx point into the pixel's tile range (0 / -1 when nothing was blended)
    tiles_x = b["tile_bounds"][0]
    rows, cols = torch.meshgrid(torch.arange(H, device=DEV), torch.arange(W, device=DEV), indexing="ij")
    rng = bins[((rows // BW) * tiles_x + cols // BW).long()]
    blended = T < 1.0
    fi, di = f["final_idx"], f["depth_idx"]
    assert bool(((fi >= rng[..., 0]) & (fi < rng[..., 1]))[blended].all()) and bool((fi[~blended] == 0).all())
    assert bool(((di == -1) | ((di >= rng[..., 0]) & (di <= fi)))[blended].all())
    # linearity in the colours and in the texture (geometry, hence every skip / stop decision, unchanged)
    g = torch.Generator().manual_seed(5)
    c2 = torch.rand(N, 3, generator=g).to(DEV)
    t2 = torch.rand(s["texture"].shape, generator=g).to(DEV)
    a_, b_ = 0.75, -1.5
    fa, _ = forward_cuda(dict(s0, colors=c2, texture=t2), ids, bins)
    fm, _ = forward_cuda(dict(s0, colors=a_ * s["colors"] + b_ * c2, texture=a_ * s["texture"] + b_ * t2), ids, bins)
    for k in ("out_img", "out_texture"):
        want = a_ * f[k] + b_ * fa[k]
        err = float((fm[k] - want).abs().max())
        print(f"  linearity {k}: max |d| = {err:.3e}")
        assert err <= 2e-5  # fp32 accumulation of ~30 terms of magnitude <= 2
    for k in ("final_Ts", "out_depth", "out_reg", "out_normal", "final_idx", "depth_idx"):
        assert torch.equal(fm[k], f[k]), k  # geometry outputs do not depend on colours / texels


def test_full_size_backward_linearity_and_adjoint(c4):
    s, b = c4
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    s0 = dict(s, background=torch.zeros(3, device=DEV))
    f, scratch = forward_cuda(s0, ids, bins)
    v1, v2 = random_vout(s0, 1), random_vout(s0, 2)
    g1 = backward_cuda(s0, ids, bins, f, v1, scratch=scratch)
    g2 = backward_cuda(s0, ids, bins, f, v2, scratch=scratch)
    v12 = {k: (v1[k] + 2.0 * v2[k]).contiguous() for k in v1}
    g12 = backward_cuda(s0, ids, bins, f, v12, scratch=scratch)
    for k in g1:
        want = g1[k] + 2.0 * g2[k]
        scale = float(want.abs().max()) + 1e-12
        bad = ((g12[k] - want).abs() > 2e-4 * scale + 2e-3 * want.abs()).float().mean()
        print(f"  backward linearity {k}: out-of-tolerance fraction {float(bad):.2e}")
        assert float(bad) < 1e-4, k  # atomic-order noise only
    # adjoint identity for the linear inputs: <v_colors, dc> + <v_texture, dt> == <v_img, J dc> + <v_tex, J dt>
    gen = torch.Generator().manual_seed(9)
    dc = torch.randn(N, 3, generator=gen).to(DEV)
    dt = torch.randn(s["texture"].shape, generator=gen).to(DEV)
    fd, _ = forward_cuda(dict(s0, colors=dc, texture=dt), ids, bins)  # J applied to (dc, dt): the forward is linear
    zero = lambda t: torch.zeros_like(t)  # noqa: E731
    v = dict(v1, v_out_depth=zero(v1["v_out_depth"]), v_out_reg=zero(v1["v_out_reg"]), v_out_alpha=zero(v1["v_out_alpha"]),
             v_out_normal=zero(v1["v_out_normal"]))
    g = backward_cuda(s0, ids, bins, f, v, scratch=scratch)
    lhs = float((g["v_colors"].double() * dc.double()).sum() + (g["v_texture"].double() * dt.double()).sum())
    rhs = float((v["v_out_img"].double() * fd["out_img"].double()).sum()
                + (v["v_out_texture"].double() * fd["out_texture"].double()).sum())
    print(f"  adjoint: <J^T v, d> = {lhs:.6e}   <v, J d> = {rhs:.6e}")
    assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs)) + 1e-3 * float(np.sqrt(W * H))
