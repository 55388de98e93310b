"""The B200 kernels' algebra (tests/formulation.py, float64) against the CPU oracle (reference algebra).

This is the CPU-side proof that the homography / moment-accumulation formulation used by the CUDA
kernels is the same function, with the same gradient, as the reference's
per-pixel quat->R / ray-plane chain.  Tolerance: the oracle is fp32, the model fp64; rtol 2e-4 /
atol 2e-5 on images, rtol 2e-3 on gradients (+1e-4 of the gradient's max magnitude).
"""
import os

import numpy as np
import pytest

import oracle
from formulation import render


def _scene(golden_dir, name, settings=None):
    g = dict(np.load(os.path.join(golden_dir, name)))
    if settings is not None:
        g["settings"] = settings
    g["H"], g["W"], g["block_width"] = int(g["H"]), int(g["W"]), int(g["block_width"])
    g["v_out"] = {k: g[k].astype(np.float64) for k in
                  ("v_out_img", "v_out_depth", "v_out_reg", "v_out_alpha", "v_out_texture", "v_out_normal")}
    return g


def _oracle(g):
    fx, fy, cx, cy = [float(v) for v in g["intrins"]]
    args = (g["H"], g["W"], g["block_width"], g["texture_dims"], g["gaussian_ids_sorted"], g["tile_bins"],
            g["colors"], g["opacities"], g["means"], g["scales"], float(g["glob_scale"]), g["quats"], g["uv0"],
            g["umap"], g["vmap"], g["texture"], g["viewmat"], g["c2w"], fx, fy, cx, cy, int(g["settings"]),
            g["background"])
    f = oracle.texture_forward(*args)
    b = oracle.texture_backward(*args, f["final_Ts"], f["final_idx"], f["depth_idx"], f["out_reg_s"],
                                g["v_out_img"], g["v_out_depth"], g["v_out_reg"], g["v_out_alpha"],
                                g["v_out_texture"], g["v_out_normal"])
    return f, b


CASES = [
    ("torch_impl_raster_c1.npz", None),
    ("torch_impl_raster_b.npz", None),
    ("torch_impl_raster_nouv.npz", None),
    ("torch_impl_raster_b.npz", (1 << 8) | (1 << 9)),            # blur
    ("torch_impl_raster_b.npz", (1 << 8) | (1 << 10)),           # ndc distortion
    ("torch_impl_raster_c1.npz", (1 << 8) | (1 << 2)),           # nearest texel
]


@pytest.mark.parametrize("name,settings", CASES)
def test_formulation_matches_oracle(golden_dir, name, settings):
    g = _scene(golden_dir, name, settings)
    f_o, b_o = _oracle(g)
    f_m, b_m = render(g, g, int(g["settings"]))
    for k in ("final_idx", "depth_idx"):
        np.testing.assert_array_equal(f_m[k], f_o[k], err_msg=k)
    for k in ("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "out_reg_s"):
        np.testing.assert_allclose(f_m[k], f_o[k], rtol=2e-4, atol=2e-5, err_msg=k)
    for k in oracle.BWD_KEYS:
        ref = b_o[k]
        atol = 1e-7 + 1e-4 * float(np.abs(ref).max())
        np.testing.assert_allclose(b_m[k].reshape(ref.shape), ref, rtol=2e-3, atol=atol, err_msg=k)
