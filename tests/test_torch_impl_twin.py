"""CPU: the package's pure-PyTorch rasteriser (gstex_cuda_b200/_torch_impl.py::texture_forward, what
``texture_gaussians(..., use_torch_impl=True)`` runs) against golden vectors produced by the reference's own
``_torch_impl.texture_forward`` + torch autograd (tests/golden/make_golden_torch_impl.py): outputs and all nine
gradients of the example.py loss on example.py's initialisation."""
import os

import numpy as np
import pytest
import torch

from gstex_cuda_b200 import _torch_impl as T


@pytest.mark.parametrize("name", ["torch_impl_raster_c1.npz", "torch_impl_raster_b.npz", "torch_impl_raster_nouv.npz"])
def test_torch_rasteriser_matches_reference_twin(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name))
    H, W, bw = int(d["H"]), int(d["W"]), int(d["block_width"])
    leaf = lambda k: torch.from_numpy(d[k]).requires_grad_(True)  # noqa: E731
    names = ("colors", "opacities", "means", "scales", "quats", "uv0", "umap", "vmap", "texture")
    p = {k: leaf(k) for k in names}
    fx, fy, cx, cy = [float(v) for v in d["intrins"]]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    n = p["means"].shape[0]
    outs = T.texture_forward(tb, (bw, bw, 1), (W, H, 1), (n, 1, 3), torch.from_numpy(d["texture_dims"]),
                             torch.from_numpy(d["gaussian_ids_sorted"]), torch.from_numpy(d["tile_bins"]), p["colors"],
                             p["opacities"], p["means"], p["scales"], float(d["glob_scale"]), p["quats"], p["uv0"], p["umap"],
                             p["vmap"], p["texture"], torch.from_numpy(d["viewmat"]), torch.from_numpy(d["c2w"]), fx, fy, cx,
                             cy, int(d["settings"]), torch.from_numpy(d["background"]))
    out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, _ = outs
    got = dict(out_img=out_img, out_depth=out_depth, out_reg=out_reg, out_alpha=1 - final_Ts, out_texture=out_texture,
               out_normal=out_normal)
    for k, v in got.items():
        torch.testing.assert_close(v.detach(), torch.from_numpy(d[k]), rtol=1e-4, atol=2e-5, msg=lambda m, k=k: f"{k}: {m}")
    # the fixture's upstream gradients are d(loss)/d(outputs) of the example.py loss
    loss = sum((got[k] * torch.from_numpy(d["v_" + k])).sum() for k in got)
    loss.backward()
    for k in names:
        ref = torch.from_numpy(d["v_" + k])
        g = p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])
        scale = float(ref.abs().max()) + 1e-12
        assert float((g - ref).abs().max()) <= 1e-3 * scale + 1e-7, f"gradient of {k}"  # fp32 summation order (matmul vs broadcast sums)


def _oracle_and_twin_gradients(shrink_uv):
    """The example.py loss on a 40-Gaussian scene: gradients from the CPU oracle (the reference CUDA kernels' semantics) and
    from autograd over the package's torch rasteriser (the reference twin's semantics), same binning."""
    import oracle
    from gstex_cuda_b200.scenes import random_small_scene

    s = random_small_scene(40, 48, 32, seed=23, device="cpu", jagged=True)
    s["opacities"] = (0.6 * s["opacities"]).contiguous()  # keeps alpha below both caps (0.99 / 0.999)
    if shrink_uv:
        s["umap"], s["vmap"] = (0.12 * s["umap"]).contiguous(), (0.12 * s["vmap"]).contiguous()
        s["uv0"] = torch.full_like(s["uv0"], 0.5)
    H, W, bw, (fx, fy, cx, cy) = s["H"], s["W"], 16, s["intrins"]
    a = {k: v.numpy() for k, v in s.items() if torch.is_tensor(v)}
    b = oracle.bin_view(a["means"], a["scales"], 1.0, a["quats"], a["viewmat"], s["intrins"], H, W, bw)
    args = (H, W, bw, a["texture_dims"], b["gaussian_ids_sorted"], b["tile_bins"], a["colors"], a["opacities"], a["means"],
            a["scales"], 1.0, a["quats"], a["uv0"], a["umap"], a["vmap"], a["texture"], a["viewmat"], a["c2w"], fx, fy, cx,
            cy, 1 << 8, a["background"])
    f = oracle.texture_forward(*args)
    v_tex = (2 * (f["out_texture"] - a["target"]) / (3 * H * W)).astype(np.float32)
    zeros = np.zeros((H, W), np.float32)
    g = oracle.texture_backward(*args, f["final_Ts"], f["final_idx"], f["depth_idx"], f["out_reg_s"],
                                np.zeros((H, W, 3), np.float32), zeros, zeros, zeros, v_tex, np.zeros((H, W, 3), np.float32))
    names = ("colors", "opacities", "means", "scales", "quats", "uv0", "umap", "vmap", "texture")
    p = {k: s[k].clone().requires_grad_(True) for k in names}
    outs = T.texture_forward(((W + bw - 1) // bw, (H + bw - 1) // bw, 1), (bw, bw, 1), (W, H, 1), (40, 1, 3),
                             s["texture_dims"], torch.from_numpy(b["gaussian_ids_sorted"]), torch.from_numpy(b["tile_bins"]),
                             p["colors"], p["opacities"], p["means"], p["scales"], 1.0, p["quats"], p["uv0"], p["umap"],
                             p["vmap"], p["texture"], s["viewmat"], s["c2w"], fx, fy, cx, cy, 1 << 8, s["background"])
    (outs[3] * torch.from_numpy(v_tex)).sum().backward()
    return g, {k: (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])).numpy() for k in names}


def test_uv_clamp_gradient_where_the_two_reference_rasterisers_differ():
    """Where u or v leaves [0, 1] the reference CUDA backward still sends the texture gradient through uv
    (texture.cu:608-609 clamps the coordinate, texture_helpers.cuh:252-300 differentiates regardless) while torch.clamp in
    its twin (_torch_impl.py:337-338) stops it.  The oracle and the kernels follow the CUDA side (pinned by the ref_cuda_*
    golden vectors); the package's torch rasteriser follows the twin.  With the uv maps kept inside the unit square the two
    agree - which is the situation example.py --torch_compare (and tests/test_gpu_api.py) is in."""
    rel = lambda got, ref: float(np.abs(got.reshape(ref.shape) - ref).max() / np.abs(ref).max())  # noqa: E731
    g, t = _oracle_and_twin_gradients(shrink_uv=True)
    for ko, kt in (("v_means", "means"), ("v_quats", "quats"), ("v_uv0", "uv0"), ("v_umap", "umap"), ("v_texture", "texture")):
        assert rel(g[ko], t[kt]) < 1e-4, ko
    g, t = _oracle_and_twin_gradients(shrink_uv=False)
    assert rel(g["v_texture"], t["texture"]) < 1e-4  # texel gradients do not go through uv
    assert rel(g["v_means"], t["means"]) > 1e-2 and rel(g["v_uv0"], t["uv0"]) > 1e-2  # the reference's own CUDA / twin gap
