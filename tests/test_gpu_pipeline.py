"""-m gpu: the fused multi-view step (gstex_cuda_b200.pipeline) against the reference-shaped API path it fuses,
and against the CPU oracle for the SH-colour stage."""
import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import sh as SH
from gstex_cuda_b200.pipeline import FusedTrainStep, DataParallelTrainStep
from gstex_cuda_b200.scenes import synthetic_scene, arc_cameras
from gstex_cuda_b200.texture import texture_gaussians
from gstex_cuda_b200.get_aabb_2d import get_aabb_2d, get_num_tiles_hit_2d, project_points
from gpu_util import DEV, to_np, assert_close_frac

pytestmark = pytest.mark.gpu

PARAMS = ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")


def _api_step(s, cams, targets):
    """The same training step through the public autograd API (one view at a time, torch glue for SH+0.5 clamp and
    the example.py loss), gradients accumulated by autograd."""
    leaves = {k: s[k].clone().requires_grad_(True) for k in PARAMS}
    H, W, bw, intr = s["H"], s["W"], 16, s["intrins"]
    total = 0.0
    for (vm, c2w), gt in zip(cams, targets):
        dirs = leaves["means"].detach() - c2w[:3, 3]
        colors = torch.clamp(SH.spherical_harmonics(s["sh_degree"], dirs, leaves["sh_coeffs"]) + 0.5, 0.0, 1.0)
        _, depths = project_points(leaves["means"].detach(), vm, intr)
        centers, extents = get_aabb_2d(leaves["means"].detach(), leaves["scales"].detach(), 1.0, leaves["quats"].detach(), vm, intr)
        nth = get_num_tiles_hit_2d(centers, extents, H, W, bw)
        # the reference counts one tile for a clipped Gaussian whose projected mean is on screen and then leaves
        # zero-filled key slots; the fused path counts 0 for them, so do the same here
        nth = torch.where((extents <= 1e-4).all(dim=-1), torch.zeros_like(nth), nth)
        outs = texture_gaussians(s["texture_info"], s["texture_dims"], centers, extents, depths, nth, colors,
                                 leaves["opacities"], leaves["means"], leaves["scales"], 1.0, leaves["quats"], leaves["uv0"],
                                 leaves["umap"], leaves["vmap"], leaves["texture"], vm, c2w, *intr, H, W, bw, 1 << 8,
                                 s["background"])
        n_ = outs[5]
        loss = (torch.nn.functional.mse_loss(outs[4], gt) + outs[2].mean()
                + (n_[..., 0] ** 2 + n_[..., 1] ** 2 + (1 - n_[..., 2]) ** 2).mean())
        loss.backward()
        total += float(loss)
    return total, {k: v.grad for k, v in leaves.items()}


@pytest.mark.parametrize("nviews", [1, 3])
def test_fused_step_matches_api_path(nviews):
    s = synthetic_scene(30000, 320, 192, seed=7, device=DEV)
    cams = [(s["viewmat"], s["c2w"])] + [(a.to(DEV), b.to(DEV)) for a, b in arc_cameras(5)[: nviews - 1]]
    g = torch.Generator().manual_seed(1)
    targets = [torch.rand(s["H"], s["W"], 3, generator=g).to(DEV) for _ in range(nviews)]
    fused = FusedTrainStep({k: s[k] for k in PARAMS}, s["texture_dims"], s["H"], s["W"], intrins=s["intrins"],
                           sh_degree=s["sh_degree"], background=s["background"], max_intersects=40 * s["num_points"])
    loss = fused.step(cams, targets)
    m = fused.check_overflow()
    assert m > 0
    assert float((1 - fused.out["final_Ts"]).mean()) > 0.03  # the last (rotated) view really rendered the scene
    with pytest.raises(RuntimeError):  # column-major camera matrices (torch.linalg.inv) are rejected, not misread
        fused.view_forward(cams[0][0], torch.linalg.inv(cams[0][0]))
    loss_api, grads = _api_step(s, cams, targets)
    assert abs(float(loss) - loss_api) <= 1e-4 * abs(loss_api) + 1e-6
    names = dict(means="v_means", scales="v_scales", quats="v_quats", opacities="v_opacity", sh_coeffs="v_sh_coeffs",
                 uv0="v_uv0", umap="v_umap", vmap="v_vmap", texture="v_texture")
    for k, gk in names.items():
        ref = to_np(grads[k])
        got = to_np(fused.grads[gk]).reshape(ref.shape)
        # same kernels on both sides: only the order of the atomic additions differs
        assert_close_frac(gk, got, ref, 1e-3, 1e-9 + 2e-5 * float(np.abs(ref).max()), 1e-4, 12, 0.05)
    # a second step on the same object reuses every buffer and gives the same result
    loss2 = float(fused.step(cams, targets))
    assert abs(loss2 - float(loss)) <= 1e-5 * abs(loss2) + 1e-7


def test_rgba_texture_layout_equals_padded_path_and_empty_shard_is_zero():
    """texture_rgba=True: the (X,4) texture is read, and its gradient accumulated, in place - no padding passes - and
    gives what the (X,3) path gives.  A step without views (a rank whose shard is empty) leaves zero gradients."""
    s = synthetic_scene(20000, 256, 160, seed=11, device=DEV)
    cams = [(s["viewmat"], s["c2w"])] + [(a.to(DEV), b.to(DEV)) for a, b in arc_cameras(5)[:2]]
    g = torch.Generator().manual_seed(2)
    targets = [torch.rand(s["H"], s["W"], 3, generator=g).to(DEV) for _ in range(3)]
    kw = dict(intrins=s["intrins"], sh_degree=s["sh_degree"], background=s["background"])
    f3 = FusedTrainStep({k: s[k] for k in PARAMS}, s["texture_dims"], s["H"], s["W"], **kw)
    p4 = {k: s[k] for k in PARAMS}
    p4["texture"] = torch.cat([s["texture"], torch.zeros_like(s["texture"][:, :1])], 1).contiguous()
    f4 = FusedTrainStep(p4, s["texture_dims"], s["H"], s["W"], texture_rgba=True, **kw)
    l3, l4 = float(f3.step(cams, targets)), float(f4.step(cams, targets))
    assert f3.check_overflow() > 0 and f4.launches < f3.launches  # no pad / un-pad kernels
    assert abs(l3 - l4) <= 1e-6 * abs(l3)
    for k in f3.grads:
        a, b = to_np(f3.grads[k]), to_np(f4.grads[k])
        if k == "v_texture":
            assert float(np.abs(b[:, 3]).max()) == 0.0
            b = b[:, :3]
        assert_close_frac(k, b.reshape(a.shape), a, 1e-3, 1e-9 + 2e-5 * float(np.abs(a).max()), 1e-4, 12, 0.05)
    f4.step([], [])
    assert float(f4.grad_arena.abs().max()) == 0.0 and float(f4.loss) == 0.0


def test_sh_colors_fused_matches_oracle():
    from gstex_cuda_b200 import _lib
    s = synthetic_scene(5000, 64, 64, seed=3, device=DEV)
    n = s["num_points"]
    colors = torch.empty((n, 3), device=DEV)
    mask = torch.empty((n,), dtype=torch.uint8, device=DEV)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.gstex_sh_colors_forward(n, 3, 3, s["means"].data_ptr(), s["c2w"].data_ptr(), s["sh_coeffs"].data_ptr(),
                                       colors.data_ptr(), mask.data_ptr(), st) == 0
    dirs = to_np(s["means"]) - to_np(s["c2w"])[:3, 3]
    raw = oracle.sh_forward(3, 3, dirs, to_np(s["sh_coeffs"])) + 0.5
    np.testing.assert_allclose(to_np(colors), np.clip(raw, 0, 1), rtol=2e-5, atol=2e-6)
    v = torch.randn(n, 3, device=DEV)
    vco = torch.empty((n, 16, 3), device=DEV)
    assert lib.gstex_sh_colors_backward(n, 3, 3, s["means"].data_ptr(), s["c2w"].data_ptr(), v.data_ptr(),
                                        mask.data_ptr(), vco.data_ptr(), 0, st) == 0
    gate = ((raw > 0) & (raw < 1)).astype(np.float32)
    want = oracle.sh_backward(3, 3, dirs, to_np(v) * gate)
    safe = np.abs(raw - np.clip(raw, 1e-5, 1 - 1e-5)) == 0  # ignore colours within rounding of the clamp
    ok = safe.all(axis=1)
    np.testing.assert_allclose(to_np(vco)[ok], want[ok], rtol=2e-5, atol=2e-6)


def test_view_sharding_is_a_partition():
    for nv, ws in ((64, 8), (7, 4), (3, 8), (64, 1)):
        parts = [DataParallelTrainStep.shard(nv, r, ws) for r in range(ws)]
        assert sorted(sum(parts, [])) == list(range(nv))
