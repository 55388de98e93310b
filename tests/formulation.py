"""float64 numpy model of the B200 rasteriser's FORMULATION (not of the reference's).

The CUDA kernels do not evaluate the reference's per-pixel quat->R / ray-plane / delta chain.  They
evaluate, per (pixel, Gaussian) pair, three affine forms in the pixel offset e = p - p_c,

    N1 = c1 + P1.e     N2 = c2 + P2.e     D = c3 + A3.e       (+ Nu, Nv for the texture coordinate)

from a per-view, per-Gaussian packed record (csrc/pack.cuh), and the backward kernel accumulates the
moments  sum g*(ex, ey, 1)  of the gradients of those forms, which a per-Gaussian epilogue
(csrc/epilogue.cuh) turns into gradients of means / scales / quats / uv maps.

This file restates exactly that algebra in numpy float64 so that the CPU test
suite can check it against the oracle without a GPU: tests/test_formulation.py.  The CUDA code
mirrors these functions line by line (same names).
"""
import math

import numpy as np

K_SIGMA = math.sqrt(0.5 * math.log2(math.e))  # alpha = opac * 2^-(l1'^2 + l2'^2)
LN2 = math.log(2.0)
T_NEAR, T_FAR = 0.01, 1000.0


def axes_from_quat(q):
    w, x, y, z = q
    a1 = np.array([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)])
    a2 = np.array([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)])
    a3 = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])
    return a1, a2, a3


def axes_vjp(q, g1, g2, g3):
    w, x, y, z = q
    return np.array([
        2 * (x * (g2[2] - g3[1]) + y * (g3[0] - g1[2]) + z * (g1[1] - g2[0])),
        2 * (-2 * x * (g2[1] + g3[2]) + y * (g1[1] + g2[0]) + z * (g1[2] + g3[0]) + w * (g2[2] - g3[1])),
        2 * (x * (g1[1] + g2[0]) - 2 * y * (g1[0] + g3[2]) + z * (g2[2] + g3[1]) + w * (g3[0] - g1[2])),
        2 * (x * (g1[2] + g3[0]) + y * (g2[2] + g3[1]) - 2 * z * (g1[0] + g2[1]) + w * (g1[1] - g2[0])),
    ])


def pack_record(mean, scale, quat, opac, uv0, umap, vmap, glob_scale, c2w, viewmat, intr):
    """csrc/pack.cuh: per-view record of one Gaussian."""
    fx, fy, cx, cy = intr
    o = c2w[:3, 3]
    Rc = c2w[:3, :3]
    a1, a2, a3 = axes_from_quat(quat)
    d = mean - o
    c0 = a3 @ d
    b1, b2, bu, bv = a1 @ d, a2 @ d, umap @ d, vmap @ d
    w1 = c0 * a1 - b1 * a3
    w2 = c0 * a2 - b2 * a3
    wu = c0 * umap - bu * a3
    wv = c0 * vmap - bv * a3
    h1, h2, hu, hv, h3 = Rc.T @ w1, Rc.T @ w2, Rc.T @ wu, Rc.T @ wv, Rc.T @ a3
    k1 = K_SIGMA / (scale[0] * glob_scale)
    k2 = K_SIGMA / (scale[1] * glob_scale)
    mc = Rc.T @ d
    exact_center = mc[2] > 1e-4 and abs(mc[0] / mc[2]) < 1e3 and abs(mc[1] / mc[2]) < 1e3
    rc = np.array([mc[0] / mc[2], mc[1] / mc[2], 1.0]) if exact_center else np.array([0.0, 0.0, 1.0])

    def form(h, kappa, vanish):
        return np.array([kappa * h[0] / fx, kappa * h[1] / fy, 0.0 if vanish else kappa * (h @ rc)])

    pv = viewmat[:3, :3] @ mean + viewmat[:3, 3]
    rw = 1.0 / (pv[2] + 1e-6)
    return dict(
        xc=fx * rc[0] + cx, yc=fy * rc[1] + cy, rc=rc, c0=c0, opac=opac,
        F1=form(h1, k1, exact_center), F2=form(h2, k2, exact_center), F3=form(h3, 1.0, False),
        FU=form(hu, 1.0, exact_center), FV=form(hv, 1.0, exact_center), uv0=np.array(uv0, dtype=np.float64),
        normal=a3, mean2d=np.array([pv[0] * rw * fx + cx, pv[1] * rw * fy + cy]), pview=pv,
        # kept for the epilogue
        a1=a1, a2=a2, a3=a3, d=d, b1=b1, b2=b2, bu=bu, bv=bv, k1=k1, k2=k2,
    )


def pixel_consts(col, row, c2w, viewmat, intr):
    fx, fy, cx, cy = intr
    px, py = col + 0.5, row + 0.5
    Rw = c2w[:3, :3] @ np.array([(px - cx) / fx, (py - cy) / fy, 1.0])
    rn = np.linalg.norm(Rw)
    ray = Rw / rn
    view_depth = viewmat[2, :3] @ ray
    return px, py, rn, view_depth


def eval_pair(rec, px, py, rn, use_blur):
    """Forward per-pair evaluation (csrc/raster_common.cuh: eval_pair)."""
    ex, ey = px - rec["xc"], py - rec["yc"]
    N1 = rec["F1"][0] * ex + rec["F1"][1] * ey + rec["F1"][2]
    N2 = rec["F2"][0] * ex + rec["F2"][1] * ey + rec["F2"][2]
    D = rec["F3"][0] * ex + rec["F3"][1] * ey + rec["F3"][2]
    eps = 1e-6 * rn
    if abs(D) < eps:
        D = eps if D >= 0 else -eps
    rD = 1.0 / D
    l1, l2 = N1 * rD, N2 * rD
    q = l1 * l1 + l2 * l2
    e = 2.0 ** (-q)
    s = rec["c0"] * rD
    t = s * rn
    bl = 0.0
    e_blur = 0.0
    if use_blur:
        dx, dy = rec["mean2d"][0] - px, rec["mean2d"][1] - py
        sb = dx * dx + dy * dy  # 0.5 * 2 * |.|^2
        e_blur = math.exp(-sb)
        if sb < q * LN2:
            bl = 1.0
    a_raw = rec["opac"] * ((1.0 - bl) * e + bl * e_blur)
    alpha = min(0.99, a_raw)
    return dict(ex=ex, ey=ey, rD=rD, l1=l1, l2=l2, e=e, s=s, t=t, alpha=alpha, bl=bl, e_blur=e_blur)


def texel_setup(dims, u, v, bilinear):
    h, w, si = int(dims[0]), int(dims[1]), int(dims[2])
    tu, tv = h * u, w * v
    i0, j0 = int(tu), int(tv)
    i1, j1 = min(i0 + 1, h - 1), min(j0 + 1, w - 1)
    fu, fv = tu - i0, tv - j0
    i0, j0 = min(i0, h - 1), min(j0, w - 1)
    wts = [(1 - fu) * (1 - fv), (1 - fu) * fv, fu * (1 - fv), fu * fv]
    idx = [si + i0 * w + j0, si + i0 * w + j1, si + i1 * w + j0, si + i1 * w + j1]
    if not bilinear:
        pick = 3
        if wts[0] >= wts[1] and wts[0] >= wts[2] and wts[0] >= wts[3]:
            pick = 0
        elif wts[1] >= wts[0] and wts[1] >= wts[2] and wts[1] >= wts[3]:
            pick = 1
        elif wts[2] >= wts[0] and wts[2] >= wts[1] and wts[2] >= wts[3]:
            pick = 2
        wts = [1.0 if k == pick else 0.0 for k in range(4)]
    return idx, wts, fu, fv, h, w


def clamp01(x):
    return min(max(x, 0.0), 1.0)


def render(scene, lists, settings):
    """Forward + backward of the formulation for a whole (small) image.  Returns (fwd, grads)."""
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    intr = tuple(float(v) for v in scene["intrins"])
    c2w, viewmat = scene["c2w"].astype(np.float64), scene["viewmat"].astype(np.float64)
    N = scene["means"].shape[0]
    C = scene["texture"].shape[1]
    g64 = lambda k: scene[k].astype(np.float64)  # noqa: E731
    means, scales, quats, opac, colors = g64("means"), g64("scales"), g64("quats"), g64("opacities").reshape(-1), g64("colors")
    uv0, umap, vmap = g64("uv0").reshape(N, 2), g64("umap").reshape(N, 3), g64("vmap").reshape(N, 3)
    texture, bg, dims = g64("texture"), g64("background"), scene["texture_dims"]
    gs = float(scene["glob_scale"])
    use_blur, use_ndc = bool(settings & (1 << 9)), bool(settings & (1 << 10))
    bilinear, prop = not (settings & (1 << 2)), bool(settings & (1 << 8))
    recs = [pack_record(means[g], scales[g], quats[g], opac[g], uv0[g], umap[g], vmap[g], gs, c2w, viewmat, intr)
            for g in range(N)]
    ids, bins = lists["gaussian_ids_sorted"], lists["tile_bins"]
    tiles_x = (W + bw - 1) // bw
    f = dict(out_img=np.zeros((H, W, 3)), out_depth=np.zeros((H, W)), out_reg=np.zeros((H, W)),
             out_texture=np.zeros((H, W, C)), out_normal=np.zeros((H, W, 3)), final_Ts=np.zeros((H, W)),
             final_idx=np.zeros((H, W), np.int32), depth_idx=np.zeros((H, W), np.int32), out_reg_s=np.zeros((H, W, 3)))
    vo = scene.get("v_out")
    acc = np.zeros((N, 32))  # per-Gaussian moment accumulators (vrec)
    v_texture = np.zeros_like(texture)

    def tex_uv(rec, pe):
        nu = rec["FU"][0] * pe["ex"] + rec["FU"][1] * pe["ey"] + rec["FU"][2]
        nv = rec["FV"][0] * pe["ex"] + rec["FV"][1] * pe["ey"] + rec["FV"][2]
        return nu * pe["rD"], nv * pe["rD"]

    for row in range(H):
        for col in range(W):
            tile = (row // bw) * tiles_x + col // bw
            lo, hi = int(bins[tile, 0]), int(bins[tile, 1])
            px, py, rn, vdep = pixel_consts(col, row, c2w, viewmat, intr)
            T, last, dlast, depth, reg = 1.0, 0, -1, 0.0, 0.0
            S = np.zeros(3)
            ac, an, at = np.zeros(3), np.zeros(3), np.zeros(C)
            for idx in range(lo, hi):
                g = int(ids[idx])
                rec = recs[g]
                pe = eval_pair(rec, px, py, rn, use_blur)
                skip = pe["t"] < T_NEAR or pe["t"] > T_FAR or pe["alpha"] < 1.0 / 255.0
                nT = T * (1 - pe["alpha"])
                if nT <= 1e-4:
                    break
                if skip:
                    continue
                vis = pe["alpha"] * T
                ac += vis * colors[g]
                an += vis * rec["normal"]
                du, dv = tex_uv(rec, pe)
                u, v = clamp01(rec["uv0"][0] + du), clamp01(rec["uv0"][1] + dv)
                tidx, wts, _, _, _, _ = texel_setup(dims[g], u, v, bilinear)
                at += vis * sum(wts[k] * texture[tidx[k]] for k in range(4))
                t_view = pe["t"] * vdep
                if T > 0.5:
                    depth, dlast = t_view, idx
                tv = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view) if use_ndc else pe["t"]
                reg += vis * (tv * tv * S[0] + S[2] - 2 * tv * S[1])
                S += vis * np.array([1.0, tv, tv * tv])
                T, last = nT, idx
            f["final_Ts"][row, col], f["final_idx"][row, col], f["depth_idx"][row, col] = T, last, dlast
            f["out_img"][row, col] = ac + T * bg
            f["out_normal"][row, col], f["out_texture"][row, col] = an, at
            f["out_depth"][row, col], f["out_reg"][row, col], f["out_reg_s"][row, col] = depth, reg, S
            if vo is None:
                continue
            # ---------------- backward replay (csrc/raster_backward.cu) ----------------
            v_img, v_n, v_tex = vo["v_out_img"][row, col], vo["v_out_normal"][row, col], vo["v_out_texture"][row, col]
            v_dep, v_reg, v_alp = vo["v_out_depth"][row, col], vo["v_out_reg"][row, col], vo["v_out_alpha"][row, col]
            v_T_run = bg @ v_img - v_alp
            for idx in range(min(last, hi - 1), lo - 1, -1):
                g = int(ids[idx])
                rec = recs[g]
                pe = eval_pair(rec, px, py, rn, use_blur)
                if pe["t"] < T_NEAR or pe["t"] > T_FAR or pe["alpha"] < 1.0 / 255.0:
                    continue
                alpha, rD, ex, ey = pe["alpha"], pe["rD"], pe["ex"], pe["ey"]
                T = T / (1 - alpha)
                vis = alpha * T
                A = acc[g]
                A[20:23] += vis * v_img            # v_rgb
                A[24:27] += vis * v_n              # direct normal term
                v_vis = colors[g] @ v_img + rec["normal"] @ v_n
                du, dv = tex_uv(rec, pe)
                u, v = clamp01(rec["uv0"][0] + du), clamp01(rec["uv0"][1] + dv)
                tidx, wts, fu, fv, th, tw = texel_setup(dims[g], u, v, bilinear)
                corner = [texture[tidx[k]] for k in range(4)]
                val = sum(wts[k] * corner[k] for k in range(4))
                v_val = vis * v_tex
                for k in range(4):
                    v_texture[tidx[k]] += wts[k] * v_val
                v_u = v_v = 0.0
                if bilinear and prop:
                    v_u = th * float(v_val @ (-(1 - fv) * corner[0] - fv * corner[1] + (1 - fv) * corner[2] + fv * corner[3]))
                    v_v = tw * float(v_val @ (-(1 - fu) * corner[0] + (1 - fu) * corner[1] - fu * corner[2] + fu * corner[3]))
                v_vis += float(val @ v_tex)
                v_alpha = T * v_vis - T * v_T_run
                v_T_cur = alpha * v_vis + (1 - alpha) * v_T_run
                t_view = pe["t"] * vdep
                t_ndc = (T_FAR * t_view - T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view)
                tv = t_ndc if use_ndc else pe["t"]
                Sf = f["out_reg_s"][row, col]
                v_tv = 2 * (vis * tv * Sf[0] - vis * Sf[1]) * v_reg
                v_w = (tv * tv * Sf[0] - 2 * tv * Sf[1] + Sf[2]) * v_reg
                v_alpha += v_w * T
                v_T_cur += v_w * alpha
                v_T_run = v_T_cur
                v_t = 0.0 if use_ndc else v_tv
                v_tview = v_dep if (idx == dlast and dlast != -1) else 0.0
                if use_ndc:
                    v_tview += (T_FAR * T_NEAR) / ((T_FAR - T_NEAR) * t_view * t_view) * v_tv
                v_t += vdep * v_tview
                v_s = v_t * rn
                # alpha = min(.99, opac*((1-bl)*2^-q + bl*e_blur)); the cap is NOT masked (texture.cu:672)
                A[7] += ((1 - pe["bl"]) * pe["e"] + pe["bl"] * pe["e_blur"]) * v_alpha      # v_opac
                v_q = -(1 - pe["bl"]) * LN2 * rec["opac"] * pe["e"] * v_alpha
                v_l1, v_l2 = 2 * pe["l1"] * v_q, 2 * pe["l2"] * v_q
                gN1, gN2 = v_l1 * rD, v_l2 * rD
                gNu, gNv = v_u * rD, v_v * rD
                gD = -(v_l1 * pe["l1"] + v_l2 * pe["l2"] + v_s * pe["s"] + v_u * du + v_v * dv) * rD
                m = np.array([ex, ey, 1.0])
                A[0:3] += gN1 * m
                A[3] += v_s * rD                    # v_c0
                A[4:7] += gN2 * m
                A[8:11] += gD * m
                A[12:15] += gNu * m
                A[15] += v_u
                A[16:19] += gNv * m
                A[19] += v_v
                if pe["bl"]:
                    v_sb = -rec["opac"] * pe["e_blur"] * v_alpha
                    A[28] += 2.0 * v_sb * (rec["mean2d"][0] - px)
                    A[29] += 2.0 * v_sb * (rec["mean2d"][1] - py)
    if vo is None:
        return f, None
    grads = epilogue(recs, acc, scales, quats, umap, vmap, gs, c2w, viewmat, intr)
    grads["v_texture"] = v_texture
    return f, grads


def epilogue(recs, acc, scales, quats, umap, vmap, glob_scale, c2w, viewmat, intr):
    """csrc/epilogue.cuh: moments -> parameter gradients, one Gaussian at a time."""
    fx, fy, cx, cy = intr
    Rc = c2w[:3, :3]
    N = len(recs)
    out = dict(v_colors=np.zeros((N, 3)), v_opacity=np.zeros((N, 1)), v_means=np.zeros((N, 3)),
               v_scales=np.zeros((N, 3)), v_quats=np.zeros((N, 4)), v_uv0=np.zeros((N, 1, 2)),
               v_umap=np.zeros((N, 1, 3)), v_vmap=np.zeros((N, 1, 3)))
    for g, rec in enumerate(recs):
        A = acc[g]
        rc = rec["rc"]

        def v_h(G, kappa):  # dL/dh from the moments of dL/dN
            return kappa * np.array([G[0] / fx + G[2] * rc[0], G[1] / fy + G[2] * rc[1], G[2]])

        G1, G2, G3, GU, GV = A[0:3], A[4:7], A[8:11], A[12:15], A[16:19]
        v_w1, v_w2 = Rc @ v_h(G1, rec["k1"]), Rc @ v_h(G2, rec["k2"])
        v_wu, v_wv = Rc @ v_h(GU, 1.0), Rc @ v_h(GV, 1.0)
        v_a3 = Rc @ v_h(G3, 1.0) + A[24:27]
        a1, a2, a3, d, c0 = rec["a1"], rec["a2"], rec["a3"], rec["d"], rec["c0"]
        v_c0 = A[3] + v_w1 @ a1 + v_w2 @ a2 + v_wu @ umap[g] + v_wv @ vmap[g]
        v_b1, v_b2, v_bu, v_bv = -(v_w1 @ a3), -(v_w2 @ a3), -(v_wu @ a3), -(v_wv @ a3)
        v_a1 = c0 * v_w1 + v_b1 * d
        v_a2 = c0 * v_w2 + v_b2 * d
        v_um = c0 * v_wu + v_bu * d
        v_vm = c0 * v_wv + v_bv * d
        v_a3 = v_a3 - rec["b1"] * v_w1 - rec["b2"] * v_w2 - rec["bu"] * v_wu - rec["bv"] * v_wv + v_c0 * d
        v_d = v_b1 * a1 + v_b2 * a2 + v_bu * umap[g] + v_bv * vmap[g] + v_c0 * a3
        # blur: gradient of the projected mean (helpers.cuh:155-164, texture.cu:685-692)
        pv = rec["pview"]
        rw = 1.0 / (pv[2] + 1e-6)
        gx, gy = fx * A[28], fy * A[29]
        v_pv = np.array([gx * rw, gy * rw, -(gx * pv[0] + gy * pv[1]) * rw * rw])
        v_d = v_d + viewmat[:3, :3].T @ v_pv
        out["v_means"][g] = v_d
        F1, F2 = rec["F1"], rec["F2"]
        out["v_scales"][g, 0] = -(F1[0] * G1[0] + F1[1] * G1[1] + F1[2] * G1[2]) / scales[g, 0]
        out["v_scales"][g, 1] = -(F2[0] * G2[0] + F2[1] * G2[1] + F2[2] * G2[2]) / scales[g, 1]
        out["v_quats"][g] = axes_vjp(quats[g], v_a1, v_a2, v_a3)
        out["v_uv0"][g, 0] = [A[15], A[19]]
        out["v_umap"][g, 0], out["v_vmap"][g, 0] = v_um, v_vm
        out["v_colors"][g], out["v_opacity"][g, 0] = A[20:23], A[7]
    return out
