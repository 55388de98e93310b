"""Generate golden vectors from the UNMODIFIED reference CUDA extension (oracle/_ref/gstex_ref_C.so, built by
oracle/build_ref.py) on a GPU box:

    gpurun -- 'python tests/golden/make_golden_ref_cuda.py'      # writes gpurun_out/golden/ref_cuda_*.npz

The files are then copied to tests/golden/ and committed; tests/test_oracle_golden_ref_cuda.py (CPU) checks the
oracle against them, which pins the oracle to the reference's own CUDA rasteriser (quirks included: alpha cap
0.99, skip / stop rules, median depth, final-sum distortion gradient).  Inputs are stored in the fixture.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gstex_cuda_b200.scenes import random_small_scene  # noqa: E402  (scene generator only: plain torch)

DEV = "cuda:0"
OUT = os.path.join(ROOT, "gpurun_out", "golden")


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so")
    spec = importlib.util.spec_from_file_location("gstex_ref_C", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def npy(t):
    return t.detach().cpu().numpy()


def make(ref, name, n, W, H, bw, settings, C, seed, opaque=False):
    s = random_small_scene(n, W, H, seed=seed, channels=C, device=DEV)
    if opaque:
        s["opacities"][:] = 1.0
    fx, fy, cx, cy = s["intrins"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    # reference pipeline, example.py:146-152 + texture.py:195-222 (torch glue restated with torch ops)
    vm = s["viewmat"]
    depths = (s["means"] @ vm[:3, :3].T + vm[:3, 3])[:, 2].contiguous()
    centers, extents = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], vm, fx, fy, cx, cy)
    tl = torch.floor((centers - extents) / bw).to(torch.int32)
    br = torch.floor((centers + extents) / bw + 1).to(torch.int32)
    tmin = torch.stack([tl[:, 0].clamp(0, tb[0]), tl[:, 1].clamp(0, tb[1])], -1)
    tmax = torch.stack([br[:, 0].clamp(0, tb[0]), br[:, 1].clamp(0, tb[1])], -1)
    nth = ((tmax - tmin)[:, 0] * (tmax - tmin)[:, 1]).to(torch.int32)
    cum = torch.cumsum(nth, 0, dtype=torch.int32)
    m = int(cum[-1])
    isect, gids = ref.map_gaussian_to_intersects(n, m, centers, extents, depths, cum, tb, bw, False)
    isect_s, perm = torch.sort(isect)
    gids_s = torch.gather(gids, 0, perm)
    bins = ref.get_tile_bin_edges(m, isect_s, tb)
    outs = ref.texture_forward(tb, (bw, bw, 1), (W, H, 1), (n, 1, C), s["texture_dims"], gids_s, bins, s["colors"],
                               s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"],
                               s["texture"], vm, s["c2w"], fx, fy, cx, cy, settings, s["background"])
    g = torch.Generator().manual_seed(seed)
    mk = lambda *shape: torch.randn(*shape, generator=g).to(DEV)  # noqa: E731
    vout = dict(v_out_img=mk(H, W, 3), v_out_depth=0.1 * mk(H, W), v_out_reg=0.1 * mk(H, W), v_out_alpha=mk(H, W),
                v_out_texture=mk(H, W, C), v_out_normal=mk(H, W, 3))
    grads = ref.texture_backward(H, W, bw, (n, 1, C), s["texture_dims"], gids_s, bins, s["colors"], s["opacities"],
                                 s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"],
                                 vm, s["c2w"], fx, fy, cx, cy, settings, s["background"], outs[5], outs[6], outs[7],
                                 outs[8], vout["v_out_img"], vout["v_out_depth"], vout["v_out_reg"], vout["v_out_alpha"],
                                 vout["v_out_texture"], vout["v_out_normal"])
    torch.cuda.synchronize()
    d = dict(H=H, W=W, block_width=bw, settings=settings, glob_scale=1.0, intrins=np.array(s["intrins"], np.float32),
             centers=npy(centers), extents=npy(extents), depths=npy(depths), num_tiles_hit=npy(nth),
             cum_tiles_hit=npy(cum), isect_ids=npy(isect), gaussian_ids=npy(gids), isect_ids_sorted=npy(isect_s),
             gaussian_ids_sorted=npy(gids_s), tile_bins=npy(bins))
    for k in ("means", "scales", "quats", "colors", "opacities", "uv0", "umap", "vmap", "texture", "texture_dims",
              "viewmat", "c2w", "background"):
        d[k] = npy(s[k])
    for k, o in zip(("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "final_idx",
                     "depth_idx", "out_reg_s"), outs):
        d[k] = npy(o)
    for k, v in vout.items():
        d[k] = npy(v)
    for k, v in zip(("v_colors", "v_opacity", "v_means", "v_scales", "v_quats", "v_uv0", "v_umap", "v_vmap", "v_texture"),
                    grads):
        d[k] = npy(v)
    # texture_edit on the same scene (texture_edit.cu:238-354): edit canvas + depth window around the rendered depth
    ge = torch.Generator().manual_seed(seed + 1000)
    upd_img = torch.rand(H, W, 3, generator=ge).to(DEV)
    upd_alpha = (torch.rand(H, W, 1, generator=ge) > 0.3).float().to(DEV) * torch.rand(H, W, 1, generator=ge).to(DEV)
    dmed = outs[1]
    zlo, zhi = (dmed - 0.5).contiguous(), (dmed + 0.5).contiguous()
    X = int(s["texture"].shape[0])
    edit_settings = (1 if settings & (1 << 9) else 0)  # edit: bit 0 = blur
    upd = ref.texture_edit(tb, (bw, bw, 1), (W, H, 1), (n, 1, 5), X, s["texture_dims"], upd_img, upd_alpha, zlo, zhi,
                           gids_s, bins, s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"],
                           s["vmap"], vm, s["c2w"], fx, fy, cx, cy, edit_settings, s["background"])
    torch.cuda.synchronize()
    d.update(edit_img=npy(upd_img), edit_alpha=npy(upd_alpha), edit_depth_lower=npy(zlo), edit_depth_upper=npy(zhi),
             edit_settings=edit_settings, edit_updated_texture=npy(upd))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "M =", m, "bytes =", os.path.getsize(os.path.join(OUT, name)))


def make_vis(ref, name, n, W, H, bw, settings, C, seed):
    """Viewer-only modes (settings bits 15-29, forward only) and wrapped (torus) key emission: SURVEY 8f rank 4."""
    s = random_small_scene(n, W, H, seed=seed, channels=C, device=DEV, spread=20.0)
    fx, fy, cx, cy = s["intrins"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    vm = s["viewmat"]
    depths = (s["means"] @ vm[:3, :3].T + vm[:3, 3])[:, 2].contiguous()
    centers, extents = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], vm, fx, fy, cx, cy)

    def bin_(wrapped):
        tc, te = centers / bw, extents / bw
        if wrapped:  # helpers.cuh:53-73 (C truncation, no clamp, boxes starting at or before tile 0 grow by one)
            lo = torch.trunc(tc - te).to(torch.int32)
            lo = torch.where(lo <= 0, lo - 1, lo)
            hi = torch.trunc(tc + te + 1).to(torch.int32)
            nth = ((hi - lo).clamp(min=0)[:, 0] * (hi - lo).clamp(min=0)[:, 1]).to(torch.int32)
            nth = torch.where((extents[:, 0] <= 1e-4) & (extents[:, 1] <= 1e-4), torch.zeros_like(nth), nth)
        else:
            tl = torch.floor(tc - te).to(torch.int32)
            br = torch.floor(tc + te + 1).to(torch.int32)
            tmin = torch.stack([tl[:, 0].clamp(0, tb[0]), tl[:, 1].clamp(0, tb[1])], -1)
            tmax = torch.stack([br[:, 0].clamp(0, tb[0]), br[:, 1].clamp(0, tb[1])], -1)
            nth = ((tmax - tmin)[:, 0] * (tmax - tmin)[:, 1]).to(torch.int32)
        cum = torch.cumsum(nth, 0, dtype=torch.int32)
        m = int(cum[-1])
        isect, gids = ref.map_gaussian_to_intersects(n, m, centers, extents, depths, cum, tb, bw, wrapped)
        isect_s, perm = torch.sort(isect)
        gids_s = torch.gather(gids, 0, perm)
        return nth, cum, m, isect, gids, isect_s, gids_s, ref.get_tile_bin_edges(m, isect_s, tb)

    nth_w, cum_w, m_w, isect_w, gids_w, isect_ws, gids_ws, bins_w = bin_(True)
    _, _, m, _, _, _, gids_s, bins = bin_(False)
    outs = ref.texture_forward(tb, (bw, bw, 1), (W, H, 1), (n, 1, C), s["texture_dims"], gids_s, bins, s["colors"],
                               s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"],
                               s["texture"], vm, s["c2w"], fx, fy, cx, cy, settings, s["background"])
    torch.cuda.synchronize()
    d = dict(H=H, W=W, block_width=bw, settings=settings, glob_scale=1.0, intrins=np.array(s["intrins"], np.float32),
             centers=npy(centers), extents=npy(extents), depths=npy(depths), gaussian_ids_sorted=npy(gids_s),
             tile_bins=npy(bins), wrapped_num_tiles_hit=npy(nth_w), wrapped_cum_tiles_hit=npy(cum_w),
             wrapped_isect_ids=npy(isect_w), wrapped_gaussian_ids=npy(gids_w), wrapped_isect_ids_sorted=npy(isect_ws),
             wrapped_gaussian_ids_sorted=npy(gids_ws), wrapped_tile_bins=npy(bins_w))
    for k in ("means", "scales", "quats", "colors", "opacities", "uv0", "umap", "vmap", "texture", "texture_dims",
              "viewmat", "c2w", "background"):
        d[k] = npy(s[k])
    for k, o in zip(("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "final_idx",
                     "depth_idx", "out_reg_s"), outs):
        d[k] = npy(o)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "M =", m, "wrapped M =", m_w, "bytes =", os.path.getsize(os.path.join(OUT, name)))


def vis_settings(alpha_bound=0.0, outline=0.0, normals=False, alpha=False, white=False, opac=False, base=1 << 8):
    return (base | (int(normals) << 15) | (int(alpha) << 16) | (int(round(alpha_bound * 8)) << 17) | (int(white) << 24)
            | (int(opac) << 25) | (int(round(outline * 4)) << 26))


if __name__ == "__main__":
    ref = load_ref()
    make_vis(ref, "refvis_cuda_normals.npz", 80, 48, 48, 16, vis_settings(normals=True), 3, 201)
    make_vis(ref, "refvis_cuda_alpha_outline.npz", 80, 64, 48, 16, vis_settings(alpha=True, alpha_bound=3.0, outline=1.0), 3, 202)
    make_vis(ref, "refvis_cuda_white_opac_c5.npz", 80, 48, 48, 8,
             vis_settings(alpha=True, alpha_bound=2.5, outline=1.5, white=True, opac=True, normals=True), 5, 203)
    make(ref, "ref_cuda_a.npz", 60, 48, 48, 16, 1 << 8, 3, 101)
    make(ref, "ref_cuda_nonsquare.npz", 150, 80, 48, 16, 1 << 8, 3, 102)
    make(ref, "ref_cuda_opaque.npz", 300, 48, 48, 16, 1 << 8, 3, 103, opaque=True)
    make(ref, "ref_cuda_blur_ndc.npz", 100, 48, 48, 16, (1 << 8) | (1 << 9) | (1 << 10), 3, 104)
    make(ref, "ref_cuda_nearest_nouv_c5.npz", 100, 48, 48, 8, (1 << 2), 5, 105)
