"""CPU: the oracle's training-step glue (preprocess forward / backward, Adam) against golden
vectors made with torch autograd over the reference's own normalized_quat_to_rotmat and with torch.optim.Adam
(tests/golden/make_golden_train_ops.py).  Tolerance: rtol 1e-5, atol 1e-6 (same fp32 formulas, different op order)."""
import os

import numpy as np

import oracle

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_ops.npz")))
RTOL, ATOL = 1e-5, 1e-6


def test_preprocess_forward_matches_reference_statements():
    o = oracle.preprocess_forward(G["raw_scales"], G["raw_quats"], G["mapping"], G["raw_rgbs"], G["raw_opacities"])
    for k in oracle.PRE_KEYS:
        np.testing.assert_allclose(o[k], G[k].reshape(o[k].shape), rtol=RTOL, atol=ATOL, err_msg=k)
    o2 = oracle.preprocess_forward(G["raw_scales"], G["raw_quats"], G["mapping"], None, G["raw_opacities"])
    assert o2["colors"] is None
    np.testing.assert_array_equal(o2["umap"], o["umap"])


def test_preprocess_backward_matches_torch_autograd():
    b = oracle.preprocess_backward(G["raw_scales"], G["raw_quats"], G["mapping"], G["raw_rgbs"], G["raw_opacities"],
                                   G["v_scales"], G["v_quats"], G["v_uv0"], G["v_umap"], G["v_vmap"], G["v_colors"],
                                   G["v_opacities"])
    for got, want in (("v_raw_scales", "g_raw_scales"), ("v_raw_quats", "g_raw_quats"), ("v_mapping", "g_mapping"),
                      ("v_raw_rgbs", "g_raw_rgbs"), ("v_raw_opacities", "g_raw_opacities")):
        ref = G[want]
        np.testing.assert_allclose(b[got].reshape(ref.shape), ref, rtol=1e-4, atol=1e-5 * float(np.abs(ref).max()),
                                   err_msg=got)


def test_texture_sigmoid_and_vjp():
    t = 1.0 / (1.0 + np.exp(-G["raw_texture"].astype(np.float32)))
    np.testing.assert_allclose(t, G["texture"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(G["v_texture"] * t * (1 - t), G["g_raw_texture"], rtol=1e-4, atol=1e-6)


def test_adam_matches_torch_optim():
    p = G["adam_p0"].copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for t in (1, 2, 3):
        oracle.adam_step(p, G[f"adam_g{t}"], m, v, 0.01, 0.9, 0.999, 1e-8, t)
        np.testing.assert_allclose(p, G[f"adam_p{t}"], rtol=RTOL, atol=ATOL, err_msg=f"step {t}")
    np.testing.assert_allclose(m, G["adam_m3"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(v, G["adam_v3"], rtol=RTOL, atol=1e-9)
