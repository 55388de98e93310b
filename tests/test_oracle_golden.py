"""Pins the CPU oracle (oracle/gstex_oracle.c) to vectors produced by the reference itself.

(a) tests/golden/torch_impl_*.npz : made by importing the reference's gstex_cuda/_torch_impl.py
    (tests/golden/make_golden_torch_impl.py).
(b) tests/golden/ref_cuda_*.npz   : made by the unmodified reference CUDA extension on a B200
    (tests/golden/make_golden_ref_cuda.py) -- see test_oracle_vs_ref_cuda.py.

Tolerances: the reference's own example asserts CUDA == _torch_impl with torch.testing.assert_close
defaults for fp32 (rtol 1.3e-6, atol 1e-5; example.py:270-275).  The oracle is a scalar C program, the
fixture a vectorised torch program, so summation order differs; we use rtol 1e-4 / atol 2e-5 on
outputs and rtol 2e-3 / atol 1e-6 + 1e-4*max|g| on gradients.
"""
import os

import numpy as np
import pytest

import oracle


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def test_sample_texture_matches_torch_impl(golden_dir):
    g = _load(golden_dir, "torch_impl_sample.npz")
    out = oracle.texture_sample_forward(g["texture_dims"], g["uvs"], g["texture"])
    np.testing.assert_allclose(out, g["out"], rtol=1.3e-6, atol=1e-5)


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh_matches_torch_impl(golden_dir, deg):
    g = _load(golden_dir, "torch_impl_sh.npz")
    colors = oracle.sh_forward(deg, deg, g["viewdirs"], g[f"coeffs{deg}"])
    np.testing.assert_allclose(colors, g[f"colors{deg}"], rtol=1.3e-6, atol=1e-5)
    vco = oracle.sh_backward(deg, deg, g["viewdirs"], g[f"v_colors{deg}"])
    np.testing.assert_allclose(vco, g[f"v_coeffs{deg}"], rtol=1.3e-6, atol=1e-5)


def test_sh_degrees_to_use_truncates(golden_dir):
    g = _load(golden_dir, "torch_impl_sh.npz")
    # evaluating a degree-4 coefficient block with degrees_to_use=2 equals the degree-2 result on the first 9 rows
    c4 = g["coeffs4"]
    got = oracle.sh_forward(4, 2, g["viewdirs"], c4)
    want = oracle.sh_forward(2, 2, g["viewdirs"], np.ascontiguousarray(c4[:, :9]))
    np.testing.assert_array_equal(got, want)
    vco = oracle.sh_backward(4, 2, g["viewdirs"], g["v_colors4"])
    assert np.all(vco[:, 9:] == 0)


def _run_oracle(g):
    fx, fy, cx, cy = [float(v) for v in g["intrins"]]
    H, W, bw, settings = int(g["H"]), int(g["W"]), int(g["block_width"]), int(g["settings"])
    args = (H, W, bw, g["texture_dims"], g["gaussian_ids_sorted"], g["tile_bins"], g["colors"], g["opacities"],
            g["means"], g["scales"], float(g["glob_scale"]), g["quats"], g["uv0"], g["umap"], g["vmap"], g["texture"],
            g["viewmat"], g["c2w"], fx, fy, cx, cy, settings, g["background"])
    f = oracle.texture_forward(*args)
    b = oracle.texture_backward(*args, f["final_Ts"], f["final_idx"], f["depth_idx"], f["out_reg_s"], g["v_out_img"],
                                g["v_out_depth"], g["v_out_reg"], g["v_out_alpha"], g["v_out_texture"],
                                g["v_out_normal"])
    return f, b


@pytest.mark.parametrize("name", ["torch_impl_raster_c1.npz", "torch_impl_raster_b.npz",
                                  "torch_impl_raster_nouv.npz"])
def test_raster_matches_torch_impl(golden_dir, name):
    g = _load(golden_dir, name)
    f, b = _run_oracle(g)
    for k in ("out_img", "out_reg", "out_texture", "out_normal", "out_depth"):
        np.testing.assert_allclose(f[k], g[k], rtol=1e-4, atol=2e-5, err_msg=k)
    np.testing.assert_allclose(1 - f["final_Ts"], g["out_alpha"], rtol=1e-4, atol=2e-5)
    for k in ("v_colors", "v_means", "v_scales", "v_quats", "v_uv0", "v_umap", "v_vmap", "v_texture"):
        ref = g[k]
        atol = 1e-6 + 1e-4 * float(np.abs(ref).max())
        np.testing.assert_allclose(b[k].reshape(ref.shape), ref, rtol=2e-3, atol=atol, err_msg=k)
    ref = g["v_opacities"]
    np.testing.assert_allclose(b["v_opacity"].reshape(ref.shape), ref, rtol=2e-3,
                               atol=1e-6 + 1e-4 * float(np.abs(ref).max()))


def test_binning_restatement_properties():
    """Tile ranges partition the sorted list; keys are sorted; ids follow emission order on ties."""
    rng = np.random.default_rng(0)
    n, H, W, bw = 500, 96, 160, 16
    centers = np.stack([rng.uniform(-20, W + 20, n), rng.uniform(-20, H + 20, n)], -1).astype(np.float32)
    extents = rng.uniform(0, 30, (n, 2)).astype(np.float32)
    depths = rng.choice(np.linspace(0.5, 9.0, 50), n).astype(np.float32)  # many ties
    nth = oracle.get_num_tiles_hit_2d(centers, extents, H, W, bw)
    m, cum = oracle.compute_cumulative_intersects(nth)
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    isect, gids, isect_s, gids_s, bins = oracle.bin_and_sort_gaussians(n, m, centers, extents, depths, cum, tb, bw)
    order = np.argsort(isect, kind="stable")
    np.testing.assert_array_equal(isect_s, isect[order])
    np.testing.assert_array_equal(gids_s, gids[order])
    tiles = (isect_s >> 32).astype(np.int64)
    for t in range(tb[0] * tb[1]):
        lo, hi = bins[t]
        assert np.all(tiles[lo:hi] == t)
        assert (hi - lo) == int((tiles == t).sum())
