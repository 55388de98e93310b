"""-m gpu: our CUDA path against the UNMODIFIED reference CUDA extension (oracle/_ref/gstex_ref_C.so, built by
oracle/build_ref.py from the sources under /root/reference) on identical inputs, on the same GPU.

Integer outputs of the binning stages must be bit-exact; float outputs use the tolerances of
test_gpu_raster.py.  The same run checks the CPU oracle against the reference CUDA (that is what pins
the oracle) and can dump golden vectors (tests/golden/make_golden_ref_cuda.py).
Skipped when the reference extension was not built.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import cuda as _C
from gstex_cuda_b200 import utils as U
from gstex_cuda_b200.scenes import random_small_scene, synthetic_scene
from gstex_cuda_b200 import sh as SH
from gpu_util import (DEV, to_np, bin_cuda, forward_cuda, backward_cuda, forward_oracle, backward_oracle, random_vout,
                      compare_forward, compare_backward, assert_close_frac)

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("reference CUDA extension not built (python oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("gstex_ref_C", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_forward(ref, s, ids, bins, bw, settings):
    H, W = s["H"], s["W"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    fx, fy, cx, cy = s["intrins"]
    outs = ref.texture_forward(tb, (bw, bw, 1), (W, H, 1), s["texture_info"], s["texture_dims"], ids, bins, s["colors"],
                               s["opacities"], s["means"], s["scales"], s["glob_scale"], s["quats"], s["uv0"], s["umap"],
                               s["vmap"], s["texture"], s["viewmat"], s["c2w"], fx, fy, cx, cy, settings, s["background"])
    torch.cuda.synchronize()
    return dict(zip(oracle.FWD_KEYS, outs))


def ref_backward(ref, s, ids, bins, bw, settings, f, vout):
    fx, fy, cx, cy = s["intrins"]
    g = ref.texture_backward(s["H"], s["W"], bw, s["texture_info"], s["texture_dims"], ids, bins, s["colors"],
                             s["opacities"], s["means"], s["scales"], s["glob_scale"], s["quats"], s["uv0"], s["umap"],
                             s["vmap"], s["texture"], s["viewmat"], s["c2w"], fx, fy, cx, cy, settings, s["background"],
                             f["final_Ts"].contiguous(), f["final_idx"].contiguous(), f["depth_idx"].contiguous(),
                             f["out_reg_s"].contiguous(), vout["v_out_img"], vout["v_out_depth"], vout["v_out_reg"],
                             vout["v_out_alpha"], vout["v_out_texture"], vout["v_out_normal"])
    torch.cuda.synchronize()
    return dict(zip(oracle.BWD_KEYS, g))


@pytest.mark.parametrize("n,W,H,bw", [(300, 96, 160, 16), (5000, 256, 192, 16), (800, 100, 60, 8)])
def test_binning_bit_exact_vs_reference_cuda(ref, n, W, H, bw):
    s = random_small_scene(n, W, H, seed=n, device=DEV)
    intr = s["intrins"]
    c_r, e_r = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], *intr)
    c_m, e_m = _C.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], *intr)
    torch.cuda.synchronize()
    assert_close_frac("centers", to_np(c_m), to_np(c_r), 1e-5, 1e-3)
    assert_close_frac("extents", to_np(e_m), to_np(e_r), 1e-5, 1e-3)
    print("  AABB bit-identical:", bool(torch.equal(c_m, c_r) and torch.equal(e_m, e_r)))
    # binning driven by the reference's own centres / extents: integers must be identical
    from gstex_cuda_b200.get_aabb_2d import get_num_tiles_hit_2d, project_points
    _, depths = project_points(s["means"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(c_r, e_r, H, W, bw)
    m, cum = U.compute_cumulative_intersects(nth)
    assert torch.equal(cum, torch.cumsum(nth, 0, dtype=torch.int32))
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    i_r, g_r = ref.map_gaussian_to_intersects(n, m, c_r, e_r, depths, cum, tb, bw, False)
    i_m, g_m = _C.map_gaussian_to_intersects(n, m, c_r, e_r, depths, cum, tb, bw, False)
    torch.cuda.synchronize()
    assert torch.equal(i_m, i_r) and torch.equal(g_m, g_r)
    is_r, perm = torch.sort(i_r)                       # utils.py:159
    gs_r = torch.gather(g_r, 0, perm)                  # utils.py:160
    is_m, gs_m = U.sort_pairs(i_m, g_m)
    assert torch.equal(is_m, is_r) and torch.equal(gs_m, gs_r)
    b_r = ref.get_tile_bin_edges(m, is_r, tb)
    b_m = _C.get_tile_bin_edges(m, is_m, tb)
    torch.cuda.synchronize()
    assert torch.equal(b_m, b_r)


@pytest.mark.parametrize("n,W,H,bw", [(300, 96, 160, 16), (3000, 256, 192, 16), (800, 100, 60, 8)])
def test_wrapped_binning_bit_exact_vs_reference_cuda(ref, n, W, H, bw):
    """wrapped=True key emission (forward.cu:34-36, 53-62): ours and the oracle against the reference kernel."""
    s = random_small_scene(n, W, H, seed=n + 5, device=DEV, spread=24.0)
    intr = s["intrins"]
    c_r, e_r = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], *intr)
    from gstex_cuda_b200.get_aabb_2d import project_points
    _, depths = project_points(s["means"], s["viewmat"], intr)
    torch.cuda.synchronize()
    nth = oracle.num_tiles_hit_wrapped(to_np(c_r), to_np(e_r), bw)
    m, cum_np = oracle.compute_cumulative_intersects(nth)
    cum = torch.from_numpy(cum_np).to(DEV)
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    i_r, g_r = ref.map_gaussian_to_intersects(n, m, c_r, e_r, depths, cum, tb, bw, True)
    i_m, g_m = _C.map_gaussian_to_intersects(n, m, c_r, e_r, depths, cum, tb, bw, True)
    torch.cuda.synchronize()
    assert torch.equal(i_m, i_r) and torch.equal(g_m, g_r)
    i_o, g_o = oracle.map_gaussian_to_intersects(n, m, to_np(c_r), to_np(e_r), to_np(depths), cum_np, tb, bw, True)
    np.testing.assert_array_equal(i_o, to_np(i_r))
    np.testing.assert_array_equal(g_o, to_np(g_r))
    assert int(((to_np(i_r) >> 32) >= tb[0] * tb[1]).sum()) == 0 and m > 0


def test_visualisation_modes_vs_reference_cuda(ref):
    """Viewer-only settings bits 15-29 (forward only): ours and the oracle against the reference kernel."""
    from test_gpu_raster import VIS_CASES, VIS_FLIP
    for settings, C, seed in VIS_CASES:
        s = random_small_scene(250, 96, 80, seed=seed, channels=C, device=DEV)
        s["settings"] = settings
        b = bin_cuda(s)
        ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
        f_m, _ = forward_cuda(s, ids, bins)
        f_r = {k: to_np(v) for k, v in ref_forward(ref, s, ids, bins, 16, settings).items()}
        print(f"settings={settings:#x}: ours vs reference CUDA")
        compare_forward(f_m, f_r, max_bad_frac=VIS_FLIP, int_bad_frac=VIS_FLIP)
        print(f"settings={settings:#x}: CPU oracle vs reference CUDA  <- pins the oracle")
        compare_forward(forward_oracle(s, to_np(ids), to_np(bins)), f_r, max_bad_frac=VIS_FLIP, int_bad_frac=VIS_FLIP)


RCASES = [
    (10, 32, 32, 16, 1 << 8, 3, 1),
    (300, 96, 64, 16, 1 << 8, 3, 2),
    (300, 64, 64, 8, 1 << 8, 3, 4),
    (400, 96, 96, 16, 0, 3, 6),
    (400, 96, 96, 16, (1 << 8) | (1 << 9), 3, 7),
    (400, 96, 96, 16, (1 << 8) | (1 << 10), 3, 8),
    (400, 96, 96, 16, (1 << 8) | (1 << 2), 3, 9),
    (300, 64, 64, 16, 1 << 8, 5, 10),
    (3000, 160, 128, 16, 1 << 8, 3, 12),
]


@pytest.mark.parametrize("n,W,H,bw,settings,C,seed", RCASES)
def test_raster_vs_reference_cuda(ref, n, W, H, bw, settings, C, seed):
    s = random_small_scene(n, W, H, seed=seed, channels=C, device=DEV)
    s["settings"], s["block_width"] = settings, bw
    b = bin_cuda(s, bw)
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    f_m, scratch = forward_cuda(s, ids, bins)
    f_r = ref_forward(ref, s, ids, bins, bw, settings)
    f_r_np = {k: to_np(v) for k, v in f_r.items()}
    print("ours vs reference CUDA (forward)")
    compare_forward(f_m, f_r_np, max_bad_frac=2e-3, int_bad_frac=2e-3)
    print("CPU oracle vs reference CUDA (forward)  <- pins the oracle")
    f_o = forward_oracle(s, to_np(ids), to_np(bins))
    compare_forward(f_o, f_r_np, max_bad_frac=2e-3, int_bad_frac=2e-3)
    vout = random_vout(s, seed)
    # all three backwards replay the REFERENCE forward's saved state
    g_r = {k: to_np(v) for k, v in ref_backward(ref, s, ids, bins, bw, settings, f_r, vout).items()}
    g_m = backward_cuda(s, ids, bins, f_r, vout, scratch=scratch)
    print("ours vs reference CUDA (backward)")
    compare_backward(g_m, g_r, max_bad_frac=2e-3)
    print("CPU oracle vs reference CUDA (backward)  <- pins the oracle")
    g_o = backward_oracle(s, to_np(ids), to_np(bins), f_r, vout)
    compare_backward(g_o, g_r, max_bad_frac=2e-3)


def test_c4_style_vs_reference_cuda(ref):
    s = synthetic_scene(100000, 640, 360, seed=1234, device=DEV)
    dirs = s["means"] - s["c2w"][:3, 3]
    col_r = ref.compute_sh_forward(s["num_points"], 3, 3, dirs.contiguous(), s["sh_coeffs"])
    col_m = _C.compute_sh_forward(s["num_points"], 3, 3, dirs.contiguous(), s["sh_coeffs"])
    torch.cuda.synchronize()
    assert_close_frac("sh colors", to_np(col_m), to_np(col_r), 1e-5, 1e-6)
    s["colors"] = torch.clamp(col_r + 0.5, 0.0, 1.0).contiguous()
    b = bin_cuda(s)
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    print("M =", b["num_intersects"])
    f_m, scratch = forward_cuda(s, ids, bins)
    f_r = ref_forward(ref, s, ids, bins, 16, 1 << 8)
    compare_forward(f_m, {k: to_np(v) for k, v in f_r.items()}, rtol=1e-3, atol=2e-4, max_bad_frac=2e-3, int_bad_frac=5e-3)
    vout = random_vout(s, 3)
    g_r = {k: to_np(v) for k, v in ref_backward(ref, s, ids, bins, 16, 1 << 8, f_r, vout).items()}
    g_r2 = {k: to_np(v) for k, v in ref_backward(ref, s, ids, bins, 16, 1 << 8, f_r, vout).items()}
    print("reference CUDA run-to-run (atomic order) jitter:")
    compare_backward(g_r2, g_r, rtol=5e-3, rel_atol=5e-4, max_bad_frac=5e-3)
    g_m = backward_cuda(s, ids, bins, f_r, vout, scratch=scratch)
    print("ours vs reference CUDA:")
    compare_backward(g_m, g_r, rtol=5e-3, rel_atol=5e-4, max_bad_frac=5e-3, outlier_bound=None)


def test_sample_and_sh_vs_reference_cuda(ref):
    g = torch.Generator().manual_seed(1)
    n = 4096
    dirs = torch.randn(n, 3, generator=g).to(DEV)
    for deg in range(5):
        K = (deg + 1) ** 2
        co = torch.rand(n, K, 3, generator=g).to(DEV)
        v = torch.randn(n, 3, generator=g).to(DEV)
        a, b = _C.compute_sh_forward(n, deg, deg, dirs, co), ref.compute_sh_forward(n, deg, deg, dirs, co)
        ga, gb = _C.compute_sh_backward(n, deg, deg, dirs, v), ref.compute_sh_backward(n, deg, deg, dirs, v)
        torch.cuda.synchronize()
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(ga, gb, rtol=2e-5, atol=2e-6)
    s = random_small_scene(50, 32, 32, seed=3, channels=7, device=DEV)
    q = 5000
    which = torch.randint(0, 50, (q,), generator=g)
    qd = s["texture_dims"][which.to(DEV)].contiguous()
    uvs = torch.rand(q, 2, generator=g).to(DEV)
    a = _C.texture_sample_forward((50, 1, 7), qd, uvs, s["texture"])
    b = ref.texture_sample_forward((50, 1, 7), qd, uvs, s["texture"])
    torch.cuda.synchronize()
    torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("yaw,pitch,roll", [(0, 0, 180), (25, 10, -30)])
def test_rotated_camera_vs_reference_cuda(ref, yaw, pitch, roll):
    """The reference's own fixtures never rotate the camera; this pins oracle and kernels for general cameras."""
    from test_gpu_raster import _rotated_scene
    s = _rotated_scene(400, 96, 96, 21, yaw, pitch, roll)
    b = bin_cuda(s)
    ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
    f_m, scratch = forward_cuda(s, ids, bins)
    f_r = ref_forward(ref, s, ids, bins, 16, 1 << 8)
    f_r_np = {k: to_np(v) for k, v in f_r.items()}
    assert float((1 - f_r_np["final_Ts"]).mean()) > 0.02
    compare_forward(f_m, f_r_np, max_bad_frac=2e-3, int_bad_frac=2e-3)
    compare_forward(forward_oracle(s, to_np(ids), to_np(bins)), f_r_np, max_bad_frac=2e-3, int_bad_frac=2e-3)
    vout = random_vout(s, 2)
    g_r = {k: to_np(v) for k, v in ref_backward(ref, s, ids, bins, 16, 1 << 8, f_r, vout).items()}
    g_m = backward_cuda(s, ids, bins, f_m, vout, scratch=scratch)
    compare_backward(g_m, g_r, max_bad_frac=2e-3)
    compare_backward(backward_oracle(s, to_np(ids), to_np(bins), f_r, vout), g_r, max_bad_frac=2e-3)
