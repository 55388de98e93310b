"""CPU: pins the oracle to vectors produced by the UNMODIFIED reference CUDA extension on a B200
(tests/golden/make_golden_ref_cuda.py).  Integer stages bit-exact; float stages with the tolerances of
tests/test_gpu_raster.py (a handful of threshold flips allowed, see there)."""
import glob
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "ref_cuda_*.npz")))


def _close(name, got, want, rtol, atol, max_bad):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    assert int(bad.sum()) <= max_bad, f"{name}: {int(bad.sum())} elements out of tolerance, max |d| {np.abs(got - want).max():.3e}"


@pytest.mark.skipif(not FILES, reason="reference-CUDA golden vectors not generated yet")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_matches_reference_cuda(path):
    g = dict(np.load(path))
    H, W, bw, settings = int(g["H"]), int(g["W"]), int(g["block_width"]), int(g["settings"])
    intr = tuple(float(v) for v in g["intrins"])
    n = g["means"].shape[0]
    # --- AABB (float) and binning (integers, from the reference's own centres / extents / depths) ---
    c, e = oracle.get_aabb_2d(g["means"], g["scales"], 1.0, g["quats"], g["viewmat"], intr)
    np.testing.assert_allclose(c, g["centers"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(e, g["extents"], rtol=1e-5, atol=1e-3)
    nth = oracle.get_num_tiles_hit_2d(g["centers"], g["extents"], H, W, bw)
    np.testing.assert_array_equal(nth, g["num_tiles_hit"])
    m, cum = oracle.compute_cumulative_intersects(nth)
    np.testing.assert_array_equal(cum, g["cum_tiles_hit"])
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    i_, g_, is_, gs_, bins = oracle.bin_and_sort_gaussians(n, m, g["centers"], g["extents"], g["depths"], cum, tb, bw)
    for got, key in ((i_, "isect_ids"), (g_, "gaussian_ids"), (is_, "isect_ids_sorted"), (gs_, "gaussian_ids_sorted"),
                     (bins, "tile_bins")):
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    # --- rasterise forward / backward ---
    fx, fy, cx, cy = intr
    args = (H, W, bw, g["texture_dims"], g["gaussian_ids_sorted"], g["tile_bins"], g["colors"], g["opacities"], g["means"],
            g["scales"], 1.0, g["quats"], g["uv0"], g["umap"], g["vmap"], g["texture"], g["viewmat"], g["c2w"], fx, fy, cx,
            cy, settings, g["background"])
    f = oracle.texture_forward(*args)
    npix = H * W
    for k in ("final_idx", "depth_idx"):
        assert int((f[k] != g[k]).sum()) <= max(2, npix // 500), k
    for k in ("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "out_reg_s"):
        _close(k, f[k], g[k], 1e-4, 2e-5, max(2, f[k].size // 500))
    b = oracle.texture_backward(*args, g["final_Ts"], g["final_idx"], g["depth_idx"], g["out_reg_s"], g["v_out_img"],
                                g["v_out_depth"], g["v_out_reg"], g["v_out_alpha"], g["v_out_texture"], g["v_out_normal"])
    for k in oracle.BWD_KEYS:
        ref = g[k]
        _close(k, b[k].reshape(ref.shape), ref, 2e-3, 1e-7 + 1e-4 * float(np.abs(ref).max()), 12)


@pytest.mark.skipif(not FILES, reason="reference-CUDA golden vectors not generated yet")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_texture_edit_matches_reference_cuda(path):
    """texture_edit: the oracle against the reference CUDA kernel's splat."""
    g = dict(np.load(path))
    if "edit_updated_texture" not in g:
        pytest.skip("fixture predates the texture_edit vectors")
    H, W, bw = int(g["H"]), int(g["W"]), int(g["block_width"])
    fx, fy, cx, cy = (float(v) for v in g["intrins"])
    want = g["edit_updated_texture"]
    got = oracle.texture_edit(H, W, bw, want.shape[1], want.shape[0], g["texture_dims"], g["edit_img"], g["edit_alpha"],
                              g["edit_depth_lower"], g["edit_depth_upper"], g["gaussian_ids_sorted"], g["tile_bins"],
                              g["opacities"], g["means"], g["scales"], 1.0, g["quats"], g["uv0"], g["umap"], g["vmap"],
                              g["viewmat"], g["c2w"], fx, fy, cx, cy, int(g["edit_settings"]))
    assert want[:, 4].sum() > 0, "fixture splats nothing"
    # a pair at the 1/255 / depth-window threshold may be kept by one side only: a few texels may differ by one splat
    _close("updated_texture", got, want, 1e-4, 1e-5, 40)


VIS_FILES = sorted(glob.glob(os.path.join(GOLDEN, "refvis_cuda_*.npz")))


@pytest.mark.skipif(not VIS_FILES, reason="reference-CUDA visualisation / wrapped golden vectors not generated yet")
@pytest.mark.parametrize("path", VIS_FILES, ids=[os.path.basename(f) for f in VIS_FILES])
def test_oracle_visualisation_and_wrapped_match_reference_cuda(path):
    """the viewer-only settings bits 15-29 (forward) and wrapped (torus) key emission."""
    g = dict(np.load(path))
    H, W, bw, settings = int(g["H"]), int(g["W"]), int(g["block_width"]), int(g["settings"])
    fx, fy, cx, cy = (float(v) for v in g["intrins"])
    n = g["means"].shape[0]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    # --- wrapped binning: integers, bit-exact ---
    nth = oracle.num_tiles_hit_wrapped(g["centers"], g["extents"], bw)
    np.testing.assert_array_equal(nth, g["wrapped_num_tiles_hit"])
    m, cum = oracle.compute_cumulative_intersects(nth)
    np.testing.assert_array_equal(cum, g["wrapped_cum_tiles_hit"])
    i_, g_, is_, gs_, bins = oracle.bin_and_sort_gaussians(n, m, g["centers"], g["extents"], g["depths"], cum, tb, bw,
                                                           wrapped=True)
    for got, key in ((i_, "isect_ids"), (g_, "gaussian_ids"), (is_, "isect_ids_sorted"), (gs_, "gaussian_ids_sorted"),
                     (bins, "tile_bins")):
        np.testing.assert_array_equal(got, g["wrapped_" + key], err_msg=key)
    # --- visualisation forward: hard footprint edges flip whole Gaussians at border pixels -> 1 % allowance ---
    f = oracle.texture_forward(H, W, bw, g["texture_dims"], g["gaussian_ids_sorted"], g["tile_bins"], g["colors"],
                               g["opacities"], g["means"], g["scales"], 1.0, g["quats"], g["uv0"], g["umap"], g["vmap"],
                               g["texture"], g["viewmat"], g["c2w"], fx, fy, cx, cy, settings, g["background"])
    assert float((g["final_Ts"] < 0.9).mean()) > 0.05, "fixture renders almost nothing"
    npix = H * W
    for k in ("final_idx", "depth_idx"):
        assert int((f[k] != g[k]).sum()) <= npix // 100, k
    for k in ("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "out_reg_s"):
        _close(k, f[k], g[k], 1e-4, 2e-5, f[k].size // 100)
