"""Helpers shared by the -m gpu parity tests: run the CUDA path through the reference-shaped backend
(gstex_cuda_b200.cuda -> C ABI) and the CPU oracle on the same inputs, and compare."""
from __future__ import annotations

import numpy as np
import torch

import oracle
from gstex_cuda_b200 import cuda as _C
from gstex_cuda_b200 import get_aabb_2d as _A
from gstex_cuda_b200 import utils as _U

DEV = "cuda:0"


def to_np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def report(name, got, want, rtol, atol):
    """Returns (max_abs_err, fraction of elements outside |d| <= atol + rtol*|want|)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    if got.size == 0:
        return 0.0, 0.0
    d = np.abs(got - want)
    bad = d > (atol + rtol * np.abs(want))
    return float(d.max()), float(bad.mean())


def assert_close_frac(name, got, want, rtol, atol, max_bad_frac=0.0, max_bad_count=0, outlier_bound=None):
    """All elements within |d| <= atol + rtol*|want|, except at most max(max_bad_frac*size, max_bad_count)
    outliers (threshold flips of the alpha / transmittance tests, see test_gpu_raster.py), each of which must
    still satisfy |d| <= outlier_bound * max|want| when a bound is given."""
    assert np.all(np.isfinite(np.asarray(got, dtype=np.float64))), f"{name}: non-finite values"
    mx, frac = report(name, got, want, rtol, atol)
    size = int(np.size(want))
    allowed = max(max_bad_frac, (max_bad_count / size) if size else 0.0)
    ref_max = float(np.abs(want).max()) if size else 0.0
    print(f"  [{name}] max|d|={mx:.3e} max|ref|={ref_max:.3e} bad_frac={frac:.2e} "
          f"(rtol={rtol}, atol={atol:.1e}, allowed {allowed:.1e})")
    assert frac <= allowed, f"{name}: {frac:.3e} of elements out of tolerance (max abs err {mx:.3e})"
    if outlier_bound is not None and size:
        assert mx <= outlier_bound * ref_max + atol, f"{name}: outlier {mx:.3e} exceeds {outlier_bound} * {ref_max:.3e}"


def bin_cuda(s, bw=None):
    """project -> AABB -> tile count -> cumsum -> emit -> sort -> ranges through the public API."""
    bw = bw or s["block_width"]
    H, W = s["H"], s["W"]
    intr = s["intrins"]
    _, depths = _A.project_points(s["means"], s["viewmat"], intr)
    centers, extents = _A.get_aabb_2d(s["means"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], intr)
    nth = _A.get_num_tiles_hit_2d(centers, extents, H, W, bw)
    m, cum = _U.compute_cumulative_intersects(nth)
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    isect, gids, isect_s, gids_s, bins = _U.bin_and_sort_gaussians(s["num_points"], m, centers, extents, depths, cum,
                                                                   tb, bw)
    return dict(depths=depths, centers=centers, extents=extents, num_tiles_hit=nth, cum_tiles_hit=cum,
                num_intersects=m, isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=isect_s,
                gaussian_ids_sorted=gids_s, tile_bins=bins, tile_bounds=tb)


def raster_args(s, ids, bins, bw=None, settings=None):
    bw = bw or s["block_width"]
    fx, fy, cx, cy = s["intrins"]
    settings = s["settings"] if settings is None else settings
    return dict(bw=bw, settings=settings, common=(s["texture_dims"], ids, bins, s["colors"], s["opacities"], s["means"],
                                                  s["scales"], s["glob_scale"], s["quats"], s["uv0"], s["umap"],
                                                  s["vmap"], s["texture"], s["viewmat"], s["c2w"], fx, fy, cx, cy,
                                                  settings, s["background"]))


def forward_cuda(s, ids, bins, bw=None, settings=None):
    a = raster_args(s, ids, bins, bw, settings)
    H, W, bw = s["H"], s["W"], a["bw"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    outs, scratch = _C.texture_forward_ex(tb, (bw, bw, 1), (W, H, 1), s["texture_info"], *a["common"])
    return dict(zip(oracle.FWD_KEYS, outs)), scratch


def backward_cuda(s, ids, bins, fwd, vout, bw=None, settings=None, scratch=None):
    a = raster_args(s, ids, bins, bw, settings)
    g = _C.texture_backward(s["H"], s["W"], a["bw"], s["texture_info"], *a["common"], fwd["final_Ts"], fwd["final_idx"],
                            fwd["depth_idx"], fwd["out_reg_s"], vout["v_out_img"], vout["v_out_depth"],
                            vout["v_out_reg"], vout["v_out_alpha"], vout["v_out_texture"], vout["v_out_normal"],
                            _fwd_scratch=scratch)
    return dict(zip(oracle.BWD_KEYS, g))


def forward_oracle(s, ids, bins, bw=None, settings=None):
    a = raster_args(s, ids, bins, bw, settings)
    return oracle.texture_forward(s["H"], s["W"], a["bw"], *[to_np(x) if torch.is_tensor(x) else x for x in a["common"]])


def backward_oracle(s, ids, bins, fwd, vout, bw=None, settings=None):
    a = raster_args(s, ids, bins, bw, settings)
    common = [to_np(x) if torch.is_tensor(x) else x for x in a["common"]]
    return oracle.texture_backward(s["H"], s["W"], a["bw"], *common, to_np(fwd["final_Ts"]), to_np(fwd["final_idx"]),
                                   to_np(fwd["depth_idx"]), to_np(fwd["out_reg_s"]), to_np(vout["v_out_img"]),
                                   to_np(vout["v_out_depth"]), to_np(vout["v_out_reg"]), to_np(vout["v_out_alpha"]),
                                   to_np(vout["v_out_texture"]), to_np(vout["v_out_normal"]))


def random_vout(s, seed=0, channels=None):
    """Random upstream gradients for all six outputs (exercises every gradient input)."""
    H, W = s["H"], s["W"]
    C = channels or s["texture"].shape[1]
    g = torch.Generator().manual_seed(seed)
    mk = lambda *shape: torch.randn(*shape, generator=g).to(DEV)  # noqa: E731
    return dict(v_out_img=mk(H, W, 3), v_out_depth=0.1 * mk(H, W), v_out_reg=0.1 * mk(H, W), v_out_alpha=mk(H, W),
                v_out_texture=mk(H, W, C), v_out_normal=mk(H, W, 3))


def compare_forward(f_c, f_o, rtol=1e-4, atol=2e-5, max_bad_frac=0.0, int_bad_frac=0.0):
    for k in ("final_idx", "depth_idx"):
        got, want = to_np(f_c[k]), f_o[k]
        frac = float((got != want).mean())
        print(f"  [{k}] mismatch fraction {frac:.2e} (allowed {int_bad_frac:.1e})")
        assert frac <= int_bad_frac, f"{k}: {frac:.3e} of pixels differ"
    for k in ("out_img", "out_reg", "out_texture", "out_normal", "final_Ts", "out_reg_s", "out_depth"):
        assert_close_frac(k, to_np(f_c[k]), f_o[k], rtol, atol, max_bad_frac)


def compare_backward(b_c, b_o, rtol=2e-3, rel_atol=1e-4, max_bad_frac=0.0, max_bad_count=12, outlier_bound=0.05):
    """A (pixel, Gaussian) pair whose alpha sits within rounding of 1/255 can be kept by one side and skipped by
    the other; that moves the <= 4 gradient entries of one Gaussian (and <= 12 texel entries) by about
    alpha*T*|v_out| <= 0.4 % of the upstream gradient.  Up to `max_bad_count` such entries per tensor are
    tolerated, none of them larger than `outlier_bound` of the tensor's largest gradient."""
    for k in oracle.BWD_KEYS:
        ref = b_o[k]
        atol = 1e-7 + rel_atol * float(np.abs(ref).max())
        assert_close_frac(k, to_np(b_c[k]).reshape(ref.shape), ref, rtol, atol, max_bad_frac, max_bad_count,
                          outlier_bound)
