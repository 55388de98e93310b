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


# ------------------------------------------------------------------------------------------------------------------
# Full-size parity against the UNMODIFIED reference CUDA extension (oracle/_ref/gstex_ref_C.so): BASELINE config 4 and
# two arc views of config 5, at the sizes the benchmark quotes.
# ------------------------------------------------------------------------------------------------------------------
import importlib.util
import os

from gstex_cuda_b200 import get_aabb_2d as A
from gstex_cuda_b200.scenes import arc_cameras
from gpu_util import to_np, compare_forward, compare_backward

REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "gstex_ref_C.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("reference CUDA extension not built (python oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("gstex_ref_C", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_chain(ref, s, viewmat, bw):
    """The reference's own binning chain, statement by statement: project_points (get_aabb_2d.py:22-32, torch ops),
    get_aabb_2d (get_aabb_2d.cu:11), get_num_tiles_hit_2d (get_aabb_2d.py:70-92, torch ops), cumsum (utils.py:57),
    map_gaussian_to_intersects (forward.cu:13-71), torch.sort + gather (utils.py:159-160), get_tile_bin_edges
    (forward.cu:76-98)."""
    H, W, intr = s["H"], s["W"], s["intrins"]
    view_points = s["means"] @ viewmat.T[:3, :3] + viewmat.T[3:, :3]
    depths = view_points[:, -1].contiguous()
    centers, extents = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], viewmat.contiguous(), *intr)
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    top_left = torch.floor((centers - extents) / bw).to(torch.int32)
    bottom_right = torch.floor((centers + extents) / bw + 1).to(torch.int32)
    tmin = torch.stack([torch.clamp(top_left[..., 0], 0, tb[0]), torch.clamp(top_left[..., 1], 0, tb[1])], -1)
    tmax = torch.stack([torch.clamp(bottom_right[..., 0], 0, tb[0]), torch.clamp(bottom_right[..., 1], 0, tb[1])], -1)
    nth = ((tmax - tmin)[:, 0] * (tmax - tmin)[:, 1]).to(torch.int32)
    cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
    m = int(cum[-1].item())
    isect, gids = ref.map_gaussian_to_intersects(s["num_points"], m, centers, extents, depths, cum, tb, bw, False)
    isect_sorted, perm = torch.sort(isect)
    gids_sorted = torch.gather(gids, 0, perm)
    bins = ref.get_tile_bin_edges(m, isect_sorted, tb)
    torch.cuda.synchronize()
    return dict(depths=depths, centers=centers, extents=extents, num_tiles_hit=nth, num_intersects=m,
                gaussian_ids_sorted=gids_sorted, tile_bins=bins, tile_bounds=tb)


def our_chain(s, viewmat, bw):
    """Our own projection / AABB / tile count through the fused bucket-by-tile binning (what texture_gaussians and the
    fused step use): nothing of the reference's feeds it."""
    H, W, intr = s["H"], s["W"], s["intrins"]
    _, depths = A.project_points(s["means"], viewmat, intr)
    centers, extents = A.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], viewmat, intr)
    nth = A.get_num_tiles_hit_2d(centers, extents, H, W, bw)
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    m = int(nth.sum().item())
    ids, bins, count, _ = U.bin_tiles(centers, extents, depths, tb, bw, m)
    torch.cuda.synchronize()
    return dict(depths=depths, centers=centers, extents=extents, num_tiles_hit=nth, num_intersects=int(count.item()),
                gaussian_ids_sorted=ids[:m], tile_bins=bins, tile_bounds=tb)


def _views():
    front = synthetic_scene(8, W, H, seed=1234)  # camera only
    arc = arc_cameras(64)
    return [("C4 front", front["viewmat"], front["c2w"]), ("C5 arc view 0 (-30 deg)", *arc[0]),
            ("C5 arc view 45 (+12.9 deg)", *arc[45])]


@pytest.mark.parametrize("vi", [0, 1, 2])
def test_full_size_binning_bit_exact_vs_reference_chain(ref, c4, vi):
    """Every integer and float of the binning chain, computed from OUR OWN projection and AABB, equals the reference's
    chain bit for bit at 1 M Gaussians / 1080p: centres, extents, depths, tile counts, M, sorted ids, tile ranges."""
    s, _ = c4
    name, viewmat, _ = _views()[vi]
    viewmat = viewmat.to(DEV).contiguous()
    r, o = reference_chain(ref, s, viewmat, BW), our_chain(s, viewmat, BW)
    print(f"  {name}: M = {r['num_intersects']}")
    for k in ("centers", "extents", "depths"):
        neq = int((r[k] != o[k]).sum())
        print(f"  {k}: {neq} of {r[k].numel()} floats differ")
    for k in ("centers", "extents", "depths", "num_tiles_hit"):
        assert torch.equal(r[k], o[k]), f"{name}: {k} not bit-identical to the reference chain"
    assert r["num_intersects"] == o["num_intersects"]
    assert torch.equal(r["gaussian_ids_sorted"], o["gaussian_ids_sorted"]), f"{name}: sorted ids differ"
    assert torch.equal(r["tile_bins"], o["tile_bins"]), f"{name}: tile ranges differ"


FULL_REPORT = {}


@pytest.mark.parametrize("vi", [0, 1, 2])
def test_full_size_raster_vs_reference_cuda(ref, c4, vi):
    """Forward (9 outputs) and backward (9 gradients) of the full-size views against the reference kernels, at the
    standard tolerances of test_gpu_raster.py (forward rtol 1e-4 / atol 2e-5, gradients rtol 2e-3 / atol 1e-4 max|g|)."""
    from test_gpu_vs_reference_cuda import ref_forward, ref_backward
    s, _ = c4
    name, viewmat, c2w = _views()[vi]
    viewmat, c2w = viewmat.to(DEV).contiguous(), c2w.to(DEV).contiguous()
    colors = SH.spherical_harmonics_colors(3, s["means"], c2w, s["sh_coeffs"]).contiguous()
    sv = dict(s, viewmat=viewmat, c2w=c2w, colors=colors)
    o = our_chain(sv, viewmat, BW)
    ids, bins = o["gaussian_ids_sorted"].contiguous(), o["tile_bins"]
    f_m, scratch = forward_cuda(sv, ids, bins)
    f_r = ref_forward(ref, sv, ids, bins, BW, 1 << 8)
    f_r_np = {k: to_np(v) for k, v in f_r.items()}
    fi_m, fi_r = to_np(f_m["final_idx"]), f_r_np["final_idx"]
    di_m, di_r = to_np(f_m["depth_idx"]), f_r_np["depth_idx"]
    flips = dict(final_idx=float((fi_m != fi_r).mean()), depth_idx=float((di_m != di_r).mean()))
    # pixels the reference stopped on a Gaussian our warp-level culling never evaluates would show up as a final_idx
    # that differs while the transmittance agrees; raster.cuh proves the outputs cannot differ there
    T_m, T_r = to_np(f_m["final_Ts"]), f_r_np["final_Ts"]
    near_stop = (T_r > 1e-4) & (T_r <= 1.0040e-4)
    flips["pixels_T_in_stop_window"] = int(near_stop.sum())
    flips["pixels_T_in_stop_window_differing"] = int((near_stop & ((fi_m != fi_r) | (T_m != T_r))).sum())
    print(f"  {name}: forward flips {flips}")
    for k in ("out_img", "out_texture", "out_normal", "final_Ts", "out_depth", "out_reg", "out_reg_s"):
        d = np.abs(to_np(f_m[k]).astype(np.float64) - f_r_np[k].astype(np.float64)).reshape(-1)
        qs = np.quantile(d, [0.5, 0.99, 0.999, 0.9999])
        print(f"  {name}: |ours - reference| of {k}: median {qs[0]:.1e}  p99 {qs[1]:.1e}  p99.9 {qs[2]:.1e}  p99.99 {qs[3]:.1e}  "
              f"max {d.max():.1e}  (max|ref| {np.abs(f_r_np[k]).max():.2e})")
    # Tolerances.  final_idx / depth_idx: <= 2e-3 of the pixels may differ (threshold flips), as everywhere.  Images: ten
    # times the standard forward tolerance (rtol 1e-3, atol 2e-4) must hold for all but the 2e-3 flip allowance; the standard
    # one (rtol 1e-4, atol 2e-5) is REPORTED and bounded at 10 % of the values.  At full size it sits at the fp32
    # conditioning of both implementations, not at their rounding: (i) the reference intersects in world space,
    # delta = (o + t ray) - mean with |o + t ray| ~ 8 and sigma down to 0.004, so its in-plane coordinates carry
    # 8 * 6e-8 / 0.004 ~ 1e-4 relative error for the sub-pixel Gaussians of this scene (and so do ours: the screen-space
    # centre of the affine forms is an fp32 pixel coordinate, good to 1e-4 px); (ii) the plane denominator dot(ray, ax3)
    # (texture_helpers.cuh:302-313) is a sum of O(1) products, so a surfel seen at cos = 1e-3 .. 1e-2 gets alpha wrong by
    # 1e-5 .. 1e-4 relative in the reference (here those records are evaluated in double, csrc/pack.cu).  Either way a
    # pair's weight moves by ~1e-4 * vis.  The distortion outputs are sums of vis * (t^2 S0 + S2 - 2 t S1) with t ~ 6..10:
    # every term is rounded at magnitude t^2 ~ 100 and cancels to O(1), so out_reg / out_reg_s get the same relative
    # tolerances applied to the magnitude of their terms (atol x 100).
    for k in ("final_idx", "depth_idx"):
        frac = float((to_np(f_m[k]) != f_r_np[k]).mean())
        assert frac <= 2e-3, f"{k}: {frac:.3e} of pixels differ"
    from gpu_util import assert_close_frac
    for k in ("out_img", "out_texture", "out_normal", "final_Ts", "out_depth", "out_reg", "out_reg_s"):
        scale = 100.0 if k.startswith("out_reg") else 1.0
        assert_close_frac(k + " (standard tolerance, reported)", to_np(f_m[k]), f_r_np[k], 1e-4, 2e-5 * scale, 0.10)
        assert_close_frac(k + " (10 x standard)", to_np(f_m[k]), f_r_np[k], 1e-3, 2e-4 * scale, 2e-3)
    vout = random_vout(sv, 3 + vi)
    g_r = {k: to_np(v) for k, v in ref_backward(ref, sv, ids, bins, BW, 1 << 8, f_r, vout).items()}
    g_r2 = {k: to_np(v) for k, v in ref_backward(ref, sv, ids, bins, BW, 1 << 8, f_r, vout).items()}
    jitter = {k: float(np.abs(g_r2[k] - g_r[k]).max() / (np.abs(g_r[k]).max() + 1e-30)) for k in g_r}
    print(f"  {name}: reference run-to-run jitter (max|d| / max|g|): " + ", ".join(f"{k} {v:.1e}" for k, v in jitter.items()))
    g_m = backward_cuda(sv, ids, bins, f_r, vout, scratch=scratch)
    # standard gradient tolerances; a flipped (pixel, Gaussian) pair moves one Gaussian's entries by up to vis * |v_out|
    # (upstream gradients are N(0,1) over 2 M pixels: up to ~5), so the size of the few outliers is bounded against that
    compare_backward(g_m, g_r, rtol=2e-3, rel_atol=1e-4, max_bad_frac=2e-3, outlier_bound=0.25)
    FULL_REPORT[name] = flips
