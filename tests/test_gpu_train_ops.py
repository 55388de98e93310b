"""-m gpu: the fused training-step glue through the C ABI.

* preprocess forward / backward, the fused texel sigmoid, Adam: against the golden vectors made with torch autograd
  over the reference's own normalized_quat_to_rotmat and with torch.optim.Adam (tests/golden/train_ops.npz), and
  against the CPU oracle on larger odd sizes.  Tolerance rtol 1e-5 / atol 1e-6 (forward, Adam), rtol 1e-4 (VJPs).
* GStexTrainStep (raw parameters -> ... -> Adam, fused) against the reference-shaped trainer: example.py:121-225 written
  with torch ops + torch autograd + torch.optim.Adam around this package's drop-in ``texture_gaussians``.
"""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from gstex_cuda_b200 import _lib
from gstex_cuda_b200.get_aabb_2d import get_aabb_2d, get_num_tiles_hit_2d, project_points
from gstex_cuda_b200.texture import texture_gaussians
from gstex_cuda_b200.trainer import GStexTrainStep
from gstex_cuda_b200 import sh as SH
from gpu_util import DEV, to_np

pytestmark = pytest.mark.gpu
G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_ops.npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def cuda_preprocess_forward(rs, rq, mp, rgb, ro):
    n = rs.shape[0]
    f = dict(dtype=torch.float32, device=DEV)
    o = dict(scales=torch.empty(n, 3, **f), quats=torch.empty(n, 4, **f), uv0=torch.empty(n, 1, 2, **f),
             umap=torch.empty(n, 1, 3, **f), vmap=torch.empty(n, 1, 3, **f),
             colors=None if rgb is None else torch.empty(n, 3, **f), opacities=torch.empty(n, 1, **f))
    P = lambda t: 0 if t is None else t.data_ptr()  # noqa: E731
    _lib.check(_lib.load().gstex_preprocess_forward(n, P(rs), P(rq), P(mp), P(rgb), P(ro), P(o["scales"]), P(o["quats"]),
                                                    P(o["uv0"]), P(o["umap"]), P(o["vmap"]), P(o["colors"]),
                                                    P(o["opacities"]), 0), "preprocess_forward")
    return o


def cuda_preprocess_backward(rs, rq, mp, rgb, ro, vs, vq, vu0, vum, vvm, vc, vo):
    n = rs.shape[0]
    f = dict(dtype=torch.float32, device=DEV)
    o = dict(v_raw_scales=torch.empty(n, 3, **f), v_raw_quats=torch.empty(n, 4, **f), v_mapping=torch.empty(n, 1, 4, **f),
             v_raw_rgbs=None if rgb is None else torch.empty(n, 3, **f), v_raw_opacities=torch.empty(n, 1, **f))
    P = lambda t: 0 if t is None else t.data_ptr()  # noqa: E731
    _lib.check(_lib.load().gstex_preprocess_backward(n, P(rs), P(rq), P(mp), P(rgb), P(ro), P(vs), P(vq), P(vu0), P(vum),
                                                     P(vvm), P(vc), P(vo), P(o["v_raw_scales"]), P(o["v_raw_quats"]),
                                                     P(o["v_mapping"]), P(o["v_raw_rgbs"]), P(o["v_raw_opacities"]), 0),
               "preprocess_backward")
    return o


def test_preprocess_forward_backward_vs_golden():
    raw = [dev(G[k]) for k in ("raw_scales", "raw_quats", "mapping", "raw_rgbs", "raw_opacities")]
    o = cuda_preprocess_forward(*raw)
    for k in oracle.PRE_KEYS:
        np.testing.assert_allclose(to_np(o[k]), G[k].reshape(o[k].shape), rtol=1e-5, atol=1e-6, err_msg=k)
    b = cuda_preprocess_backward(*raw, *[dev(G[k]) for k in ("v_scales", "v_quats", "v_uv0", "v_umap", "v_vmap",
                                                            "v_colors", "v_opacities")])
    for got, want in (("v_raw_scales", "g_raw_scales"), ("v_raw_quats", "g_raw_quats"), ("v_mapping", "g_mapping"),
                      ("v_raw_rgbs", "g_raw_rgbs"), ("v_raw_opacities", "g_raw_opacities")):
        ref = G[want]
        np.testing.assert_allclose(to_np(b[got]).reshape(ref.shape), ref, rtol=1e-4,
                                   atol=1e-5 * float(np.abs(ref).max()), err_msg=got)


@pytest.mark.parametrize("n,with_rgb", [(1, True), (1001, True), (4097, False)])
def test_preprocess_vs_oracle(n, with_rgb):
    g = torch.Generator().manual_seed(n)
    rs, rq = torch.randn(n, 3, generator=g) - 3, torch.randn(n, 4, generator=g)
    mp = torch.randn(n, 1, 4, generator=g)
    rgb = torch.randn(n, 3, generator=g) if with_rgb else None
    ro = torch.randn(n, 1, generator=g)
    ups = [torch.randn(n, *s, generator=g) for s in ((3,), (4,), (1, 2), (1, 3), (1, 3), (3,), (1,))]
    D = lambda t: None if t is None else t.to(DEV)  # noqa: E731
    N = lambda t: None if t is None else t.numpy()  # noqa: E731
    o = cuda_preprocess_forward(D(rs), D(rq), D(mp), D(rgb), D(ro))
    oo = oracle.preprocess_forward(N(rs), N(rq), N(mp), N(rgb), N(ro))
    for k in oracle.PRE_KEYS:
        if oo[k] is not None:
            np.testing.assert_allclose(to_np(o[k]), oo[k], rtol=1e-5, atol=1e-6, err_msg=k)
    if not with_rgb:
        ups[5] = None
    b = cuda_preprocess_backward(D(rs), D(rq), D(mp), D(rgb), D(ro), *[D(u) for u in ups])
    bo = oracle.preprocess_backward(N(rs), N(rq), N(mp), N(rgb), N(ro), *[N(u) for u in ups])
    for k in oracle.PRE_GRAD_KEYS:
        if bo[k] is not None:
            np.testing.assert_allclose(to_np(b[k]), bo[k], rtol=1e-4, atol=1e-5 * float(np.abs(bo[k]).max()), err_msg=k)


def test_sigmoid_pad_unpad():
    X = 5003
    g = torch.Generator().manual_seed(1)
    raw = (3 * torch.randn(X, 3, generator=g)).to(DEV)
    tex4, g4 = torch.empty(X, 4, device=DEV), torch.randn(X, 4, generator=g).to(DEV)
    lib = _lib.load()
    _lib.check(lib.gstex_sigmoid_pad_texture(X, raw.data_ptr(), tex4.data_ptr(), 0), "sigmoid_pad")
    t = torch.sigmoid(raw)
    np.testing.assert_allclose(to_np(tex4[:, :3]), to_np(t), rtol=1e-5, atol=1e-6)
    assert float(tex4[:, 3].abs().max()) == 0.0
    out = torch.full((X, 3), 7.0, device=DEV)
    _lib.check(lib.gstex_unpad_texture_grad_sigmoid(X, g4.data_ptr(), tex4.data_ptr(), out.data_ptr(), 0, 0), "unpad")
    want = g4[:, :3] * t * (1 - t)
    np.testing.assert_allclose(to_np(out), to_np(want), rtol=1e-4, atol=1e-6)
    _lib.check(lib.gstex_unpad_texture_grad_sigmoid(X, g4.data_ptr(), tex4.data_ptr(), out.data_ptr(), 1, 0), "unpad")
    np.testing.assert_allclose(to_np(out), to_np(2 * want), rtol=1e-4, atol=1e-6)


def cuda_adam(p, g, m, v, lr, b1, b2, eps, step, scale=1.0):
    _lib.check(_lib.load().gstex_adam_step(p.numel(), p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), lr, b1, b2,
                                           eps, step, scale, 0), "adam_step")


def test_adam_vs_torch_golden_and_oracle():
    p = dev(G["adam_p0"].copy())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in (1, 2, 3):
        cuda_adam(p, dev(G[f"adam_g{t}"]), m, v, 0.01, 0.9, 0.999, 1e-8, t)
        np.testing.assert_allclose(to_np(p), G[f"adam_p{t}"], rtol=1e-5, atol=1e-6, err_msg=f"step {t}")
    np.testing.assert_allclose(to_np(m), G["adam_m3"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(to_np(v), G["adam_v3"], rtol=1e-5, atol=1e-9)
    # large, odd-sized, unaligned start (scalar path) against the oracle, with a gradient scale
    gen = torch.Generator().manual_seed(5)
    n = 1_000_003
    base = torch.randn(n + 1, generator=gen)
    pp, gg = base[1:].clone(), torch.randn(n, generator=gen)
    p_o, m_o, v_o = pp.numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    buf = torch.zeros(n + 1, device=DEV)
    buf[1:] = pp.to(DEV)
    p_c = buf[1:]  # 4-byte aligned only
    m_c, v_c, g_c = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), gg.to(DEV)
    for t in (1, 2):
        oracle.adam_step(p_o, gg.numpy(), m_o, v_o, 0.003, 0.8, 0.99, 1e-6, t, 0.5)
        cuda_adam(p_c, g_c, m_c, v_c, 0.003, 0.8, 0.99, 1e-6, t, 0.5)
    np.testing.assert_allclose(to_np(p_c), p_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(to_np(v_c), v_o, rtol=1e-5, atol=1e-9)


# ----------------------------------------------------------------------------------------------------------------
def example_raw(N, H, W, th, tw, seed, sh_degree=None):
    """example.py:69-119 initialisation (raw leaves), optionally with SH coefficients instead of rgbs."""
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(N, 3, generator=g) - 0.5
    means[:, :2] *= 16
    means[:, 0] *= W / max(W, H)
    means[:, 1] *= H / max(W, H)
    scales = 0.5 * math.log(1 / N) * torch.rand(N, 3, generator=g)
    u, v, w = torch.rand(N, 1, generator=g), torch.rand(N, 1, generator=g), torch.rand(N, 1, generator=g)
    quats = torch.cat([torch.sqrt(1 - u) * torch.sin(2 * math.pi * v), torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
                       torch.sqrt(u) * torch.sin(2 * math.pi * w), torch.sqrt(u) * torch.cos(2 * math.pi * w)], -1)
    quats = quats * (0.5 + torch.rand(N, 1, generator=g))  # un-normalised leaves
    mapping = torch.zeros(N, 1, 4)
    mapping[:, :, :2] = 0.5
    mapping[:, :, 2] = math.log(0.25 * math.sqrt(1 / float(torch.sum(torch.exp(scales[:, 0] + scales[:, 1])))))
    mapping[:, :, 3] = 2 * math.pi * torch.rand(N, 1, generator=g)
    raw = dict(means=means, scales=scales, quats=quats, opacities=torch.randn(N, 1, generator=g) + 1.0,
               mapping=mapping, texture=torch.randn(N * th * tw, 3, generator=g))
    if sh_degree is None:
        raw["rgbs"] = torch.rand(N, 3, generator=g)
    else:
        K = (sh_degree + 1) ** 2
        sh = 0.1 * torch.randn(N, K, 3, generator=g)
        sh[:, 0] = (torch.rand(N, 3, generator=g) - 0.5) / 0.28209479177387814
        raw["sh_coeffs"] = sh
    dims = torch.zeros(N, 3, dtype=torch.int32)
    dims[:, 0], dims[:, 1] = th, tw
    dims[:, 2] = torch.arange(N, dtype=torch.int32) * th * tw
    f = 0.5 * W / math.tan(0.25 * math.pi)
    return {k: t.to(DEV).contiguous() for k, t in raw.items()}, dims.to(DEV), (f, f, W / 2.0, H / 2.0)


def quat_axes12(q):
    w, x, y, z = q.unbind(-1)
    a1 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], -1)
    a2 = torch.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], -1)
    return a1, a2


def reference_shaped_loss(leaves, dims, intr, H, W, cams, targets, sh_degree, bg):
    """example.py:121-209 with torch ops + autograd around the drop-in API; summed over the views."""
    scales = torch.zeros_like(leaves["scales"])
    scales[:, :2] = torch.exp(leaves["scales"][:, :2])
    scales[:, -1] = 1e-5 * torch.mean(scales[:, :-1], dim=-1).detach()
    quats = leaves["quats"] / leaves["quats"].norm(dim=-1, keepdim=True)
    a1, a2 = quat_axes12(quats)
    mp = leaves["mapping"]
    uv0, us, th = mp[:, :, :2], torch.exp(mp[:, :, None, 2]), mp[:, :, None, 3]
    umap = us * (a1[:, None] * torch.cos(th) + a2[:, None] * torch.sin(th))
    vmap = us * (-a1[:, None] * torch.sin(th) + a2[:, None] * torch.cos(th))
    N = scales.shape[0]
    total = 0.0
    for (vm, c2w), gt in zip(cams, targets):
        if "rgbs" in leaves:
            colors = torch.sigmoid(leaves["rgbs"])
        else:
            dirs = leaves["means"].detach() - c2w[:3, 3]
            colors = torch.clamp(SH.spherical_harmonics(sh_degree, dirs, leaves["sh_coeffs"]) + 0.5, 0.0, 1.0)
        with torch.no_grad():
            _, depths = project_points(leaves["means"], vm, intr)
            centers, extents = get_aabb_2d(leaves["means"], scales, 1.0, quats, vm, intr)
            nth = get_num_tiles_hit_2d(centers, extents, H, W, 16)
        outs = texture_gaussians((N, 1, 3), dims, centers, extents, depths, nth, colors, torch.sigmoid(leaves["opacities"]),
                                 leaves["means"], scales, 1.0, quats, uv0, umap, vmap, torch.sigmoid(leaves["texture"]),
                                 vm, c2w, *intr, H, W, 16, 1 << 8, bg)
        n_ = outs[5]
        total = total + (torch.nn.functional.mse_loss(outs[4], gt) + outs[2].mean()
                         + (n_[..., 0] ** 2 + n_[..., 1] ** 2 + (1 - n_[..., 2]) ** 2).mean())
    return total


@pytest.mark.parametrize("mode,views,N,H,W", [("rgb", 1, 200, 64, 64), ("sh", 2, 301, 48, 80)])
def test_fused_trainer_matches_reference_shaped_trainer(mode, views, N, H, W):
    sh_degree = 3
    raw, dims, intr = example_raw(N, H, W, 4, 3, seed=N, sh_degree=None if mode == "rgb" else sh_degree)
    vm = torch.eye(4, device=DEV)
    vm[2, 3] = 8.0
    cams = []
    for k in range(views):
        v = vm.clone()
        v[0, 3] = 0.4 * k
        cams.append((v.contiguous(), torch.linalg.inv(v).contiguous()))
    g = torch.Generator().manual_seed(9)
    targets = [torch.rand(H, W, 3, generator=g).to(DEV) for _ in range(views)]
    bg = torch.zeros(3, device=DEV)

    fused = GStexTrainStep(raw, dims, H, W, intrins=intr, sh_degree=sh_degree, lr=0.01, background=bg,
                           max_intersects=64 * N)
    leaves = {k: t.clone().requires_grad_(k != "mapping") for k, t in raw.items()}
    opt = torch.optim.Adam([t for k, t in leaves.items() if k != "mapping"], lr=0.01)

    for it in range(3):
        opt.zero_grad()
        loss_ref = reference_shaped_loss(leaves, dims, intr, H, W, cams, targets, sh_degree, bg)
        loss_ref.backward()
        loss_fused = fused.forward_backward(cams, targets)
        torch.cuda.synchronize()
        assert abs(float(loss_fused) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref)) + 1e-6, (it, float(loss_fused), float(loss_ref))
        for k, t in leaves.items():
            if t.grad is None:
                continue
            ref, got = to_np(t.grad), to_np(fused.raw_grads[k])
            atol = 1e-7 + 2e-4 * float(np.abs(ref).max())
            bad = np.abs(got - ref) > atol + 2e-3 * np.abs(ref)
            assert bad.mean() <= 2e-3, f"iteration {it}: gradient of {k}: {bad.mean():.2e} of entries off"
        opt.step()
        fused.optimizer_step()
        for k, t in leaves.items():
            got, ref = to_np(fused.raw[k]), to_np(t.detach())
            # Adam's first steps move every parameter by ~lr regardless of gradient size: compare with an lr-sized atol
            bad = np.abs(got - ref) > 2e-4 + 1e-4 * np.abs(ref)
            assert bad.mean() <= 5e-3, f"iteration {it}: parameter {k}: {bad.mean():.2e} of entries off"
    fused.fused.check_overflow()
    assert float(np.abs(to_np(fused.raw["mapping"]) - to_np(raw["mapping"])).max()) == 0.0  # frozen (example.py:118)


def test_cuda_graph_replay_matches_eager_steps():
    """GStexTrainStep.capture()/replay(): one CUDA graph per optimiser step (device-resident Adam step counter, no host
    sync anywhere in the step) against the same steps launched eagerly; prints the per-step time of both on the
    example.py default configuration (BASELINE config 2: 256x256, 100 Gaussians, 101x101 texels each)."""
    import time
    N, H, W = 100, 256, 256
    raw, dims, intr = example_raw(N, H, W, 101, 101, seed=3)
    vm = torch.eye(4, device=DEV)
    vm[2, 3] = 8.0
    cams = [(vm.contiguous(), torch.linalg.inv(vm).contiguous())]
    targets = [torch.rand(H, W, 3, generator=torch.Generator().manual_seed(4)).to(DEV)]
    bg = torch.zeros(3, device=DEV)
    mk = lambda: GStexTrainStep(raw, dims, H, W, intrins=intr, sh_degree=3, lr=0.01, background=bg, max_intersects=64 * N)  # noqa: E731
    eager, graphed = mk(), mk()
    graphed.capture(cams, targets)
    # capture() does not train: parameters and optimiser state are where they started
    assert torch.equal(graphed.param_arena, eager.param_arena) and graphed.step_count == 0
    assert float(graphed.exp_avg.abs().max()) == 0.0
    losses_e, losses_g = [], []
    for _ in range(5):
        losses_e.append(float(eager.step(cams, targets)))
        losses_g.append(float(graphed.replay()))
    assert graphed.step_count == eager.step_count == 5
    assert losses_e[-1] < losses_e[0]  # it trains
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-3 * abs(a) + 1e-6, (losses_e, losses_g)
    for k in eager.raw:
        got, ref = to_np(graphed.raw[k]), to_np(eager.raw[k])
        bad = np.abs(got - ref) > 2e-4 + 1e-4 * np.abs(ref)  # atomic-order noise through Adam (see the test above)
        assert bad.mean() <= 5e-3, f"parameter {k}: {bad.mean():.2e} of entries differ between graph replay and eager"

    def per_step(fn, iters=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return 1e6 * (time.perf_counter() - t0) / iters

    t_e, t_g = per_step(lambda: eager.step(cams, targets)), per_step(graphed.replay)
    print(f"  BASELINE config 2 optimiser step: eager {t_e:.0f} us, CUDA graph replay {t_g:.0f} us "
          f"({eager.total_launches // max(1, eager.step_count)} launches -> 1)")
    assert t_g < t_e


@pytest.mark.parametrize("jagged", [False, True])
def test_visible_only_adam_leaves_unseen_gaussians_untouched(jagged):
    """SURVEY 8f rank 3, "sparse / visible-only update": with visible_only=True the optimiser touches only Gaussians that
    hit a tile in some view of the step.  Step 1 sees every Gaussian; step 2 looks at the right half of the scene only.
    A dense Adam keeps moving the unseen Gaussians on their momentum in step 2; the visible-only one leaves their
    parameters and both moments bit-for-bit alone, and updates the seen ones exactly like the dense one."""
    N, H, W, th, tw = 400, 96, 128, 3, 2
    raw, dims, intr = example_raw(N, H, W, th, tw, seed=13)
    if jagged:  # a permuted texel layout: the texel -> Gaussian table path
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(1)).to(DEV)
        dims = dims.clone()
        dims[:, 2] = (perm * th * tw).to(torch.int32)
    vm1 = torch.eye(4, device=DEV)
    vm1[2, 3] = 8.0
    vm2 = vm1.clone()
    vm2[0, 3] = -14.0  # scene shifted far left in view space: only its right-hand part stays on screen
    cams1 = [(vm1.contiguous(), torch.linalg.inv(vm1).contiguous())]
    cams2 = [(vm2.contiguous(), torch.linalg.inv(vm2).contiguous())]
    targets = [torch.rand(H, W, 3, generator=torch.Generator().manual_seed(5)).to(DEV)]
    bg = torch.zeros(3, device=DEV)
    mk = lambda vis: GStexTrainStep(raw, dims, H, W, intrins=intr, lr=0.01, background=bg, visible_only=vis)  # noqa: E731
    sparse, dense = mk(True), mk(False)
    for t in (sparse, dense):
        t.step(cams1, targets)
    seen1 = sparse.raw_grads["visible"].clone()
    assert float((seen1 > 0).float().mean()) > 0.9
    after1 = {k: v.clone() for k, v in sparse.raw.items()}
    m1, v1 = sparse.exp_avg.clone(), sparse.exp_avg_sq.clone()
    for t in (sparse, dense):
        t.step(cams2, targets)
    torch.cuda.synchronize()
    seen2 = sparse.raw_grads["visible"] > 0
    frac = float(seen2.float().mean())
    print(f"  visible in step 2: {frac:.2f} of the Gaussians")
    assert 0.05 < frac < 0.8
    unseen = ~seen2
    moved_dense = 0.0
    for k in ("means", "scales", "quats", "opacities", "rgbs"):
        a, b, d = sparse.raw[k], after1[k], dense.raw[k]
        assert torch.equal(a[unseen], b[unseen]), f"{k}: an unseen Gaussian was updated"
        moved_dense = max(moved_dense, float((d[unseen] - b[unseen]).abs().max()))
        bad = (a[seen2] - d[seen2]).abs() > 2e-4 + 1e-4 * d[seen2].abs()
        assert float(bad.float().mean()) <= 5e-3, f"{k}: seen Gaussians differ from the dense update"
    assert moved_dense > 1e-4  # the dense optimiser did move them (momentum): the gate is what kept them still
    # texels and both moments of unseen Gaussians
    tex_owner = torch.empty(N * th * tw, dtype=torch.long, device=DEV)
    for_g = (dims[:, 2].long()[:, None] + torch.arange(th * tw, device=DEV)[None, :])
    tex_owner[for_g.reshape(-1)] = torch.arange(N, device=DEV).repeat_interleave(th * tw)
    tex_unseen = unseen[tex_owner]
    assert torch.equal(sparse.raw["texture"][tex_unseen], after1["texture"][tex_unseen])
    off = sparse.raw["means"].data_ptr() - sparse.param_arena.data_ptr()
    sl = slice(off // 4, off // 4 + 3 * N)
    assert torch.equal(sparse.exp_avg[sl].view(N, 3)[unseen], m1[sl].view(N, 3)[unseen])
    assert torch.equal(sparse.exp_avg_sq[sl].view(N, 3)[unseen], v1[sl].view(N, 3)[unseen])
    assert sparse.step_count == dense.step_count == 2
