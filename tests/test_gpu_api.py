"""-m gpu: the public autograd API (texture_gaussians and friends) end to end, example.py style."""
import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200.texture import texture_gaussians
from gstex_cuda_b200.get_aabb_2d import get_aabb_2d, get_num_tiles_hit_2d, project_points, get_aabb_2d_torch
from gstex_cuda_b200._torch_impl import normalized_quat_to_rotmat
from gstex_cuda_b200.scenes import random_small_scene
from gpu_util import DEV, to_np, assert_close_frac

pytestmark = pytest.mark.gpu


def _loss(outputs, gt):
    """example.py:189-209"""
    out_texture, out_reg, out_normal = outputs[4][:, :, :3], outputs[2], outputs[5]
    return (torch.nn.functional.mse_loss(out_texture, gt) + out_reg.mean()
            + (out_normal[:, :, 0] ** 2 + out_normal[:, :, 1] ** 2 + (1 - out_normal[:, :, 2]) ** 2).mean())


def test_example_style_training_step_matches_oracle():
    s = random_small_scene(120, 96, 80, seed=8, device=DEV)
    H, W, bw = s["H"], s["W"], 16
    leaves = {k: s[k].clone().requires_grad_(True) for k in
              ("colors", "opacities", "means", "scales", "quats", "uv0", "umap", "vmap", "texture")}
    intr = s["intrins"]
    _, depths = project_points(leaves["means"], s["viewmat"], intr)
    centers, extents = get_aabb_2d(leaves["means"], leaves["scales"], 1, leaves["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, H, W, bw)
    outs = texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, leaves["colors"],
                             leaves["opacities"], leaves["means"], leaves["scales"], 1, leaves["quats"], leaves["uv0"],
                             leaves["umap"], leaves["vmap"], leaves["texture"], s["viewmat"], s["c2w"], *intr, H, W, bw,
                             1 << 8, s["background"])
    assert [tuple(o.shape) for o in outs] == [(H, W, 3), (H, W), (H, W), (H, W), (H, W, 3), (H, W, 3)]
    loss = _loss(outs, s["target"])
    loss.backward()

    # oracle: same pipeline on the CPU, gradients of the same loss injected by hand
    npz = {k: to_np(v) for k, v in s.items() if torch.is_tensor(v)}
    b = oracle.bin_view(npz["means"], npz["scales"], 1.0, npz["quats"], npz["viewmat"], intr, H, W, bw)
    fx, fy, cx, cy = intr
    args = (H, W, bw, npz["texture_dims"], b["gaussian_ids_sorted"], b["tile_bins"], npz["colors"], npz["opacities"],
            npz["means"], npz["scales"], 1.0, npz["quats"], npz["uv0"], npz["umap"], npz["vmap"], npz["texture"],
            npz["viewmat"], npz["c2w"], fx, fy, cx, cy, 1 << 8, npz["background"])
    f = oracle.texture_forward(*args)
    for k, o in zip(("out_img", "out_depth", "out_reg"), outs[:3]):
        assert_close_frac(k, to_np(o), f[k], 1e-4, 2e-5, 2e-3)
    assert_close_frac("out_alpha", to_np(outs[3]), 1 - f["final_Ts"], 1e-4, 2e-5, 2e-3)
    P = H * W
    gt = npz["target"]
    v_tex = 2 * (f["out_texture"] - gt) / (3 * P)
    n_ = f["out_normal"]
    v_n = np.stack([2 * n_[..., 0], 2 * n_[..., 1], -2 * (1 - n_[..., 2])], -1) / P
    zeros = np.zeros((H, W), np.float32)
    g = oracle.texture_backward(*args, f["final_Ts"], f["final_idx"], f["depth_idx"], f["out_reg_s"],
                                np.zeros((H, W, 3), np.float32), zeros, np.full((H, W), 1.0 / P, np.float32), zeros,
                                v_tex.astype(np.float32), v_n.astype(np.float32))
    for k_leaf, k_o in (("colors", "v_colors"), ("opacities", "v_opacity"), ("means", "v_means"), ("scales", "v_scales"),
                        ("quats", "v_quats"), ("uv0", "v_uv0"), ("umap", "v_umap"), ("vmap", "v_vmap"),
                        ("texture", "v_texture")):
        ref = g[k_o]
        got = to_np(leaves[k_leaf].grad).reshape(ref.shape)
        assert_close_frac(k_o, got, ref, 2e-3, 1e-9 + 2e-4 * float(np.abs(ref).max()), 2e-3, 12, 0.05)


def test_no_intersections_returns_background():
    s = random_small_scene(5, 32, 32, seed=2, device=DEV)
    s["means"][:, 0] = 1000.0  # far outside the frustum: AABBs miss every tile
    intr = s["intrins"]
    means = s["means"].clone().requires_grad_(True)
    _, depths = project_points(means, s["viewmat"], intr)
    centers, extents = get_aabb_2d(means, s["scales"], 1, s["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, 32, 32, 16)
    assert int(nth.sum()) == 0
    outs = texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, s["colors"],
                             s["opacities"], means, s["scales"], 1, s["quats"], s["uv0"], s["umap"], s["vmap"],
                             s["texture"], s["viewmat"], s["c2w"], *intr, 32, 32, 16, 1 << 8, s["background"])
    assert torch.allclose(outs[0], s["background"].expand(32, 32, 3))
    outs[0].sum().backward()
    assert float(means.grad.abs().max()) == 0.0


def test_uint8_colors_and_default_background():
    s = random_small_scene(30, 48, 48, seed=3, device=DEV)
    intr = s["intrins"]
    _, depths = project_points(s["means"], s["viewmat"], intr)
    centers, extents = get_aabb_2d(s["means"], s["scales"], 1, s["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, 48, 48, 16)
    c8 = (s["colors"] * 255).to(torch.uint8)
    a = texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, c8, s["opacities"],
                          s["means"], s["scales"], 1, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"],
                          s["viewmat"], s["c2w"], *intr, 48, 48, 16, 1 << 8)
    b = texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, c8.float() / 255,
                          s["opacities"], s["means"], s["scales"], 1, s["quats"], s["uv0"], s["umap"], s["vmap"],
                          s["texture"], s["viewmat"], s["c2w"], *intr, 48, 48, 16, 1 << 8,
                          torch.ones(3, device=DEV))
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    with pytest.raises(AssertionError):
        texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, c8, s["opacities"],
                          s["means"], s["scales"], 1, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"],
                          s["viewmat"], s["c2w"], *intr, 48, 48, 32, 1 << 8)


def test_project_points_autograd_and_aabb_torch_twin():
    s = random_small_scene(200, 64, 64, seed=4, device=DEV)
    intr = s["intrins"]
    m = s["means"].clone().requires_grad_(True)
    pix, depths = project_points(m, s["viewmat"], intr)
    (pix.sum() * 0.01 + (depths ** 2).sum()).backward()
    m2 = s["means"].clone().requires_grad_(True)
    pv = m2 @ s["viewmat"][:3, :3].T + s["viewmat"][:3, 3]
    rw = 1.0 / (pv[:, 2] + 1e-6)
    pix2 = torch.stack([pv[:, 0] * rw * intr[0] + intr[2], pv[:, 1] * rw * intr[1] + intr[3]], -1)
    (pix2.sum() * 0.01 + (pv[:, 2] ** 2).sum()).backward()
    torch.testing.assert_close(pix, pix2, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(m.grad, m2.grad, rtol=1e-4, atol=1e-5)
    c, e = get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], intr)
    ct, et = get_aabb_2d_torch(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], intr)
    torch.testing.assert_close(c, ct, rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(e, et, rtol=1e-4, atol=1e-2)
    R = normalized_quat_to_rotmat(s["quats"])
    torch.testing.assert_close(R @ R.transpose(-1, -2), torch.eye(3, device=DEV).expand_as(R), rtol=1e-4, atol=1e-5)


def test_clipped_gaussians_reproduce_reference_zero_slots():
    """Reference quirk: a clipped Gaussian (mean behind the near plane) whose projected mean is on
    screen counts one tile in get_num_tiles_hit_2d but is skipped by the key emitter, leaving zero-filled slots
    (key 0 -> tile 0, Gaussian 0).  The drop-in API path reproduces that bit for bit."""
    s = random_small_scene(6, 32, 32, seed=2, device=DEV)
    s["means"][3:, 2] = -8.0 + 0.005  # z_view = 0.005 <= 0.01: clipped
    s["means"][3:, :2] = 0.0           # projects onto the principal point -> on screen
    intr = s["intrins"]
    _, depths = project_points(s["means"], s["viewmat"], intr)
    centers, extents = get_aabb_2d(s["means"], s["scales"], 1, s["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, 32, 32, 16)
    assert (to_np(extents)[3:] == 0).all() and (to_np(nth)[3:] == 1).all()
    from gstex_cuda_b200.utils import compute_cumulative_intersects, bin_and_sort_gaussians
    m, cum = compute_cumulative_intersects(nth)
    isect, gids, isect_s, gids_s, bins = bin_and_sort_gaussians(6, m, centers, extents, depths, cum, (2, 2, 1), 16)
    b = oracle.bin_and_sort_gaussians(6, m, to_np(centers), to_np(extents), to_np(depths), to_np(cum), (2, 2, 1), 16)
    for got, want in zip((isect, gids, isect_s, gids_s, bins), b):
        np.testing.assert_array_equal(to_np(got), want)
    assert int((to_np(isect) == 0).sum()) == 3


def test_view_prefetcher_ring_delivers_views_in_order():
    """gstex_cuda_b200.prefetch: double-buffered pinned-host -> device staging on a copy stream."""
    from gstex_cuda_b200.prefetch import ViewPrefetcher
    g = torch.Generator().manual_seed(0)
    views = [(torch.rand(4, 4, generator=g).pin_memory(), torch.rand(64, 48, 3, generator=g).pin_memory())
             for _ in range(5)]
    pf = ViewPrefetcher(torch.device(DEV), depth=2)
    pf.submit(views[0])
    acc = []
    for k in range(5):
        cam, img = pf.get()
        if k + 1 < 5:
            pf.submit(views[k + 1])
        acc.append((cam.clone(), (img * 2.0).sum()))  # a consumer on the compute stream
        pf.release()
    torch.cuda.synchronize()
    for k in range(5):
        assert torch.equal(acc[k][0].cpu(), views[k][0])
        torch.testing.assert_close(acc[k][1].cpu(), (views[k][1] * 2.0).sum(), rtol=1e-5, atol=1e-3)
    assert pf.bytes_copied == sum(a.numel() * 4 + b.numel() * 4 for a, b in views)
    pf.submit(views[0]); pf.submit(views[1])
    with pytest.raises(RuntimeError, match="full"):
        pf.submit(views[2])
    with pytest.raises(RuntimeError, match="pinned"):
        ViewPrefetcher(torch.device(DEV)).submit((torch.zeros(4),))


def test_fused_image_loss_and_sh_colors_match_torch_glue():
    """gstex_cuda_b200.loss.image_loss and sh.spherical_harmonics_colors against the torch ops they replace
    (example.py:189-209 and clamp(SH + 0.5, 0, 1)), values and gradients, fp32 tolerances."""
    from gstex_cuda_b200.loss import image_loss
    from gstex_cuda_b200 import sh as SH
    g = torch.Generator().manual_seed(3)
    H, W = 90, 70
    tex = torch.rand(H, W, 3, generator=g).to(DEV).requires_grad_(True)
    reg = torch.rand(H, W, generator=g).to(DEV).requires_grad_(True)
    nrm = torch.randn(H, W, 3, generator=g).to(DEV).requires_grad_(True)
    gt = torch.rand(H, W, 3, generator=g).to(DEV)
    ref = (torch.nn.functional.mse_loss(tex, gt) + reg.mean()
           + (nrm[..., 0] ** 2 + nrm[..., 1] ** 2 + (1 - nrm[..., 2]) ** 2).mean())
    g_ref = torch.autograd.grad(ref * 1.7, (tex, reg, nrm))
    got = image_loss(tex, reg, nrm, gt)
    g_got = torch.autograd.grad(got * 1.7, (tex, reg, nrm))
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    for a, b in zip(g_got, g_ref):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-9)

    n = 3000
    means = torch.randn(n, 3, generator=g).to(DEV)
    c2w = torch.eye(4); c2w[:3, 3] = torch.tensor([0.3, -0.2, -8.0]); c2w = c2w.to(DEV)
    for deg in (0, 1, 2, 3):
        K = (deg + 1) ** 2
        co = (torch.randn(n, K, 3, generator=g) * 2.5).to(DEV).requires_grad_(True)
        v = torch.randn(n, 3, generator=g).to(DEV)
        ref = torch.clamp(SH.spherical_harmonics(deg, means - c2w[:3, 3], co) + 0.5, 0.0, 1.0)
        got = SH.spherical_harmonics_colors(deg, means, c2w, co)
        assert float(((ref == 0) | (ref == 1)).float().mean()) > 0.05  # the clamp is exercised
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=2e-6)
        (g_ref,) = torch.autograd.grad(ref, co, v)
        (g_got,) = torch.autograd.grad(got, co, v)
        # a channel within rounding of the clamp can be gated on one side only
        bad = ((g_got - g_ref).abs() > 1e-5 + 1e-4 * g_ref.abs()).float().mean()
        assert float(bad) < 1e-3


def _run_api(s, leaves, **kw):
    H, W, bw, intr = s["H"], s["W"], 16, s["intrins"]
    _, depths = project_points(leaves["means"], s["viewmat"], intr)
    centers, extents = get_aabb_2d(leaves["means"], leaves["scales"], 1, leaves["quats"], s["viewmat"], intr)
    nth = get_num_tiles_hit_2d(centers, extents, H, W, bw)
    return texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, leaves["colors"],
                             leaves["opacities"], leaves["means"], leaves["scales"], 1, leaves["quats"], leaves["uv0"],
                             leaves["umap"], leaves["vmap"], leaves["texture"], s["viewmat"], s["c2w"], *intr, H, W, bw,
                             1 << 8, s["background"], **kw)


_LEAVES = ("colors", "opacities", "means", "scales", "quats", "uv0", "umap", "vmap", "texture")


def test_capacity_mode_equals_the_synchronising_call():
    """texture_gaussians(..., max_intersects=cap): no host read-back of the intersection count; same images, same
    gradients as the reference-shaped call, and last_intersect_count() reports what a too-small capacity dropped."""
    from gstex_cuda_b200 import texture as TX
    s = random_small_scene(300, 128, 96, seed=17, device=DEV)
    res = []
    for kw in ({}, {"max_intersects": 50000}):
        leaves = {k: s[k].clone().requires_grad_(True) for k in _LEAVES}
        outs = _run_api(s, leaves, **kw)
        _loss(outs, s["target"]).backward()
        res.append((outs, leaves))
    count = TX.last_intersect_count(DEV)
    assert 0 < count <= 50000
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for k in _LEAVES:
        ga, gb = res[0][1][k].grad, res[1][1][k].grad
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max()) + 1e-12, k  # atomic order only
    leaves = {k: s[k].clone() for k in _LEAVES}
    _run_api(s, leaves, max_intersects=count // 2)  # deliberately too small: detectable, no crash
    assert TX.last_intersect_count(DEV) == count > count // 2


def test_use_torch_impl_agrees_with_the_kernels_like_example_torch_compare():
    """use_torch_impl=True runs the package's pure-PyTorch rasteriser over the same binning (reference texture.py:117,
    :411-514).  On a case without alpha-cap or stop-rule activity (where the torch twin's semantics coincide with the
    kernels', as in example.py --torch_compare) outputs and autograd gradients match the CUDA path.
    The uv maps are shrunk so that every blended pixel samples strictly inside the unit square: the reference's two
    rasterisers themselves disagree where u or v is clamped - its CUDA backward keeps the uv gradient there
    (texture.cu:608-609 clamps, texture_helpers.cuh:252-300 differentiates regardless), torch.clamp in its twin
    (_torch_impl.py:337-338) zeroes it - and the kernels follow the CUDA one (tests/test_oracle_golden.py pins the
    oracle on both sides of that line: test_uv_clamp_gradient_follows_the_reference_cuda)."""
    s = random_small_scene(40, 48, 32, seed=23, device=DEV, jagged=True)
    s["opacities"] = (0.6 * s["opacities"]).contiguous()
    s["umap"], s["vmap"] = (0.12 * s["umap"]).contiguous(), (0.12 * s["vmap"]).contiguous()
    s["uv0"] = torch.full_like(s["uv0"], 0.5)
    res = []
    for flag in (False, True):
        leaves = {k: s[k].clone().requires_grad_(True) for k in _LEAVES}
        outs = _run_api(s, leaves, use_torch_impl=flag)
        _loss(outs, s["target"]).backward()
        res.append((outs, leaves))
    for k, a, b in zip(("out_img", "out_depth", "out_reg", "out_alpha", "out_texture", "out_normal"), res[0][0], res[1][0]):
        assert_close_frac(k, to_np(a), to_np(b), 1e-4, 2e-5, 5e-3)
    for k in _LEAVES:
        gr = res[1][1][k].grad  # torch autograd leaves the gradient of an input the loss does not reach as None
        ref = to_np(gr) if gr is not None else np.zeros(tuple(res[1][1][k].shape), np.float32)
        got = to_np(res[0][1][k].grad)
        assert_close_frac("grad " + k, got, ref, 2e-3, 1e-9 + 2e-4 * float(np.abs(ref).max()), 5e-3, 12, 0.1)


def test_rgba_texture_layout_through_the_autograd_api():
    """texture (X,4) with texture_info[2] == 3: the same images and gradients as the (X,3) texture, without the padding
    passes; the gradient comes back (X,4) with a zero fourth column."""
    s = random_small_scene(200, 96, 64, seed=29, device=DEV)
    res = []
    for rgba in (False, True):
        leaves = {k: s[k].clone().requires_grad_(True) for k in _LEAVES}
        if rgba:
            leaves["texture"] = torch.cat([s["texture"], torch.zeros_like(s["texture"][:, :1])], 1).contiguous().requires_grad_(True)
        outs = _run_api(s, leaves)
        _loss(outs, s["target"]).backward()
        res.append((outs, leaves))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    g3, g4 = res[0][1]["texture"].grad, res[1][1]["texture"].grad
    assert g4.shape == (s["texture"].shape[0], 4) and float(g4[:, 3].abs().max()) == 0.0
    assert float((g4[:, :3] - g3).abs().max()) <= 1e-5 * float(g3.abs().max()) + 1e-12
    for k in _LEAVES[:-1]:
        ga, gb = res[0][1][k].grad, res[1][1][k].grad
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max()) + 1e-12, k


@pytest.mark.parametrize("rgba", [False, True])
def test_in_place_gradient_accumulation_equals_autograd(rgba):
    """texture_gaussians(..., texture_grad=buf) and spherical_harmonics_colors(..., coeffs_grad=buf): two views accumulated
    straight into caller-owned buffers equal autograd's own accumulation over the same two views, and the leaves' .grad of
    the fused inputs stays untouched."""
    from gstex_cuda_b200 import sh as SH
    s = random_small_scene(150, 96, 64, seed=31, device=DEV)
    g = torch.Generator().manual_seed(5)
    coeffs0 = (torch.randn(150, 16, 3, generator=g) * 0.5).to(DEV)
    tex0 = torch.cat([s["texture"], torch.zeros_like(s["texture"][:, :1])], 1).contiguous() if rgba else s["texture"]
    views = [s["c2w"], s["c2w"].clone()]
    views[1][0, 3] += 0.4  # SH colours see a second camera position
    res = []
    for fused in (False, True):
        leaves = {k: s[k].clone().requires_grad_(True) for k in _LEAVES if k not in ("colors", "texture")}
        leaves["texture"] = tex0.clone().requires_grad_(True)
        coeffs = coeffs0.clone().requires_grad_(True)
        tg, cg = (torch.zeros_like(tex0), torch.zeros_like(coeffs0)) if fused else (None, None)
        for c2w in views:
            leaves["colors"] = SH.spherical_harmonics_colors(3, leaves["means"], c2w, coeffs, coeffs_grad=cg)
            kw = dict(texture_grad=tg) if fused else {}
            outs = _run_api(s, leaves, **kw)
            (_loss(outs, s["target"]) + (outs[0] - s["target"]).square().mean()).backward()  # colours (SH) take part
        if fused:
            assert leaves["texture"].grad is None and coeffs.grad is None
            res.append((tg, cg, leaves))
        else:
            res.append((leaves["texture"].grad, coeffs.grad, leaves))
    for a, b in zip(res[0][:2], res[1][:2]):
        assert float(a.abs().max()) > 0
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-12  # atomic order only
    for k in ("means", "scales", "quats", "opacities", "uv0", "umap", "vmap"):
        ga, gb = res[0][2][k].grad, res[1][2][k].grad
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max()) + 1e-12, k
    with pytest.raises(ValueError):
        _run_api(s, res[0][2], texture_grad=torch.zeros(3, device=DEV))
