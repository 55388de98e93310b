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
