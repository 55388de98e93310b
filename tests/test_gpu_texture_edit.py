"""-m gpu: texture_edit (reference texture_edit.py:14-239, texture_edit.cu:11-354) through the
public API and the C ABI, against the CPU oracle and - when it was built - the UNMODIFIED reference CUDA extension.

Tolerance: a texel's five sums accumulate bilinear weights of (u, v) coordinates that the two implementations compute
with algebraically equal but differently rounded fp32 formulas (|du| ~ 1e-6, so a weight moves by ~h * 1e-6), hence
`|d| <= 1e-4 * |want| + 2e-5 * (1 + total weight splatted on the texel)`.  A (pixel, Gaussian) pair whose alpha sits
within rounding of 1/255 or whose depth sits on the edge of the window can be splatted by one implementation only, so
up to 0.1 % of the texel entries may differ, each by at most one splat (<= 1 per channel)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import cuda as _C
from gstex_cuda_b200.get_aabb_2d import get_aabb_2d, get_num_tiles_hit_2d, project_points
from gstex_cuda_b200.scenes import random_small_scene
from gstex_cuda_b200.texture_edit import texture_edit
from gpu_util import DEV, to_np, bin_cuda, forward_cuda

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so")


def edit_inputs(s, f, seed, window=0.5):
    H, W = s["H"], s["W"]
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(H, W, 3, generator=g).to(DEV)
    alpha = ((torch.rand(H, W, 1, generator=g) > 0.3).float() * torch.rand(H, W, 1, generator=g)).to(DEV)
    d = f["out_depth"]
    return img, alpha, (d - window).contiguous(), (d + window).contiguous()


def assert_edit_close(got, want, max_bad_frac=1e-3, max_bad_count=20):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape and np.all(np.isfinite(got))
    tol = 1e-4 * np.abs(want) + 2e-5 * (1.0 + want[:, 4:5])
    d = np.abs(got - want)
    bad = d > tol
    allowed = max(max_bad_frac, max_bad_count / want.size)
    print(f"  [updated_texture] max|d|={d.max():.3e} max|ref|={np.abs(want).max():.3e} bad_frac={bad.mean():.2e} "
          f"(allowed {allowed:.1e})")
    assert bad.mean() <= allowed, f"{bad.mean():.3e} of texel entries out of tolerance"
    assert d.max() <= 1.0 + 1e-3, "an outlier is at most one splat"


def run_cuda(s, b, img, alpha, zlo, zhi, settings, C=5):
    H, W, bw = s["H"], s["W"], s["block_width"]
    fx, fy, cx, cy = s["intrins"]
    X = s["texture"].shape[0]
    return _C.texture_edit(b["tile_bounds"], (bw, bw, 1), (W, H, 1), (s["num_points"], 1, C), X, s["texture_dims"], img,
                           alpha, zlo, zhi, b["gaussian_ids_sorted"], b["tile_bins"], s["opacities"], s["means"],
                           s["scales"], s["glob_scale"], s["quats"], s["uv0"], s["umap"], s["vmap"], s["viewmat"],
                           s["c2w"], fx, fy, cx, cy, settings, s["background"])


def run_oracle(s, b, img, alpha, zlo, zhi, settings, C=5):
    fx, fy, cx, cy = s["intrins"]
    n = to_np
    return oracle.texture_edit(s["H"], s["W"], s["block_width"], C, s["texture"].shape[0], n(s["texture_dims"]), n(img),
                               n(alpha), n(zlo), n(zhi), n(b["gaussian_ids_sorted"]), n(b["tile_bins"]),
                               n(s["opacities"]), n(s["means"]), n(s["scales"]), s["glob_scale"], n(s["quats"]),
                               n(s["uv0"]), n(s["umap"]), n(s["vmap"]), n(s["viewmat"]), n(s["c2w"]), fx, fy, cx, cy,
                               settings)


@pytest.mark.parametrize("n,W,H,bw,settings,C", [(200, 96, 64, 16, 0, 5), (500, 100, 60, 8, 1, 5), (60, 48, 48, 16, 2, 7),
                                                 (2000, 160, 128, 16, 0, 5)])
def test_texture_edit_vs_oracle(n, W, H, bw, settings, C):
    s = random_small_scene(n, W, H, seed=n + 7, device=DEV)
    s["block_width"] = bw
    b = bin_cuda(s)
    f, _ = forward_cuda(s, b["gaussian_ids_sorted"], b["tile_bins"])
    img, alpha, zlo, zhi = edit_inputs(s, f, seed=n)
    got = run_cuda(s, b, img, alpha, zlo, zhi, settings, C)
    want = run_oracle(s, b, img, alpha, zlo, zhi, settings, C)
    torch.cuda.synchronize()
    assert got.shape == (s["texture"].shape[0], C) and want[:, 4].sum() > 0
    assert_edit_close(to_np(got), want)
    if C > 5:
        assert float(got[:, 5:].abs().max()) == 0.0


def test_texture_edit_public_api_and_depth_window():
    """The reference-shaped call (binning inside), an all-pass and an all-reject depth window."""
    s = random_small_scene(300, 96, 64, seed=11, device=DEV)
    H, W, bw, intr = s["H"], s["W"], s["block_width"], s["intrins"]
    _, depths = project_points(s["means"], s["viewmat"], intr)
    centers, extents = get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, H, W, bw)
    g = torch.Generator().manual_seed(3)
    img = torch.rand(H, W, 3, generator=g).to(DEV)
    alpha = torch.rand(H, W, 1, generator=g).to(DEV)
    zero, big = torch.zeros(H, W, device=DEV), torch.full((H, W), 1e9, device=DEV)
    args = ((300, 1, 5), s["texture_dims"], img, alpha)
    rest = (centers, extents, depths, nth, s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"],
            s["vmap"], s["viewmat"], s["c2w"], *intr, H, W, bw, 0, s["background"])
    full = texture_edit(*args, zero, big, *rest)
    none = texture_edit(*args, big, big + 1, *rest)
    assert float(none.abs().max()) == 0.0
    b = bin_cuda(s)
    want = run_oracle(s, b, img, alpha, zero, big, 0)
    assert_edit_close(to_np(full), want)
    # channel 3 <= channel 4 (alpha <= 1), rgb*a <= a
    t = to_np(full)
    assert np.all(t[:, 3] <= t[:, 4] + 1e-4) and np.all(t[:, :3] <= t[:, 3:4] + 1e-4)


def test_texture_edit_rejects_bad_arguments():
    s = random_small_scene(50, 48, 48, seed=5, device=DEV)
    b = bin_cuda(s)
    f, _ = forward_cuda(s, b["gaussian_ids_sorted"], b["tile_bins"])
    img, alpha, zlo, zhi = edit_inputs(s, f, seed=1)
    with pytest.raises(RuntimeError):
        run_cuda(s, b, img, alpha, zlo, zhi, 0, C=3)      # fewer than 5 output channels
    with pytest.raises(RuntimeError):
        run_cuda(s, b, img, alpha, zlo, zhi, 1 << 9)      # rasteriser bit positions are not edit bits
    with pytest.raises(RuntimeError):
        run_cuda(s, b, img.cpu(), alpha, zlo, zhi, 0)


@pytest.mark.parametrize("n,W,H,bw,settings", [(300, 96, 160, 16, 0), (1500, 128, 96, 16, 1), (800, 100, 60, 8, 0)])
def test_texture_edit_vs_reference_cuda(n, W, H, bw, settings):
    if not os.path.exists(REF_SO):
        pytest.skip("reference CUDA extension not built (python oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("gstex_ref_C", REF_SO)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    s = random_small_scene(n, W, H, seed=n + 3, device=DEV)
    s["block_width"] = bw
    b = bin_cuda(s)
    f, _ = forward_cuda(s, b["gaussian_ids_sorted"], b["tile_bins"])
    img, alpha, zlo, zhi = edit_inputs(s, f, seed=n + 1)
    fx, fy, cx, cy = s["intrins"]
    X = s["texture"].shape[0]
    want = ref.texture_edit(b["tile_bounds"], (bw, bw, 1), (W, H, 1), (n, 1, 5), X, s["texture_dims"], img, alpha, zlo,
                            zhi, b["gaussian_ids_sorted"], b["tile_bins"], s["opacities"], s["means"], s["scales"], 1.0,
                            s["quats"], s["uv0"], s["umap"], s["vmap"], s["viewmat"], s["c2w"], fx, fy, cx, cy, settings,
                            s["background"])
    got = run_cuda(s, b, img, alpha, zlo, zhi, settings)
    torch.cuda.synchronize()
    assert float(want[:, 4].sum()) > 0
    assert_edit_close(to_np(got), to_np(want))
