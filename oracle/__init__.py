"""CPU oracle for the textured-2DGS hot path.  TEST INFRASTRUCTURE ONLY.

ctypes/numpy front end for ``gstex_oracle.c`` (a plain-C restatement of the reference
algorithms, each function citing the reference file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package; the product (``gstex_cuda_b200``) never does.

All functions take and return numpy arrays (float32 / int32 / int64, C-contiguous).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Tuple

import numpy as np

from .build import build as _build

_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_build())
        _LIB.orc_cumsum.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(int(n)))


# ----------------------------------------------------------------------------------------------
def project_points(means, viewmat, intrins) -> Tuple[np.ndarray, np.ndarray]:
    means, viewmat = _f(means), _f(viewmat)
    n = means.shape[0]
    pix = np.empty((n, 2), np.float32)
    depths = np.empty((n,), np.float32)
    fx, fy, cx, cy = [float(v) for v in intrins]
    lib().orc_project_points(C.c_int(n), _p(means), _p(viewmat), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                             C.c_float(cy), _p(pix), _p(depths))
    return pix, depths


def get_aabb_2d(means, scales, glob_scale, quats, viewmat, intrins):
    means, scales, quats, viewmat = _f(means), _f(scales), _f(quats), _f(viewmat)
    n = means.shape[0]
    centers = np.zeros((n, 2), np.float32)
    extents = np.zeros((n, 2), np.float32)
    fx, fy, cx, cy = [float(v) for v in intrins]
    lib().orc_get_aabb_2d(C.c_int(n), _p(means), _p(scales), C.c_float(glob_scale), _p(quats), _p(viewmat),
                          C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(centers), _p(extents))
    return centers, extents


def get_num_tiles_hit_2d(centers, extents, img_height, img_width, block_width) -> np.ndarray:
    centers, extents = _f(centers), _f(extents)
    n = centers.shape[0]
    out = np.zeros((n,), np.int32)
    lib().orc_num_tiles_hit(C.c_int(n), _p(centers), _p(extents), C.c_int(img_height), C.c_int(img_width),
                            C.c_int(block_width), _p(out))
    return out


def compute_cumulative_intersects(num_tiles_hit) -> Tuple[int, np.ndarray]:
    v = _i(num_tiles_hit)
    out = np.zeros_like(v)
    total = lib().orc_cumsum(C.c_int(v.shape[0]), _p(v), _p(out)) if v.shape[0] else 0
    return int(total), out


def num_tiles_hit_wrapped(centers, extents, block_width) -> np.ndarray:
    """Tiles a Gaussian writes under wrapped=True (helpers.cuh:53-73): what cum_tiles_hit must sum."""
    centers, extents = _f(centers), _f(extents)
    out = np.zeros((centers.shape[0],), np.int32)
    lib().orc_num_tiles_hit_wrapped(C.c_int(centers.shape[0]), _p(centers), _p(extents), C.c_int(block_width), _p(out))
    return out


def map_gaussian_to_intersects(num_points, num_intersects, centers, extents, depths, cum_tiles_hit, tile_bounds,
                               block_width, wrapped=False):
    centers, extents, depths, cum = _f(centers), _f(extents), _f(depths), _i(cum_tiles_hit)
    isect = np.zeros((num_intersects,), np.int64)
    gids = np.zeros((num_intersects,), np.int32)
    lib().orc_map_gaussian_to_intersects(C.c_int(num_points), _p(centers), _p(extents), _p(depths), _p(cum),
                                         C.c_int(tile_bounds[0]), C.c_int(tile_bounds[1]), C.c_int(block_width),
                                         C.c_int(1 if wrapped else 0), _p(isect), _p(gids))
    return isect, gids


def sort_pairs(keys, vals):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    vals = _i(vals)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    lib().orc_sort_pairs(C.c_int64(keys.shape[0]), _p(keys), _p(vals), _p(ko), _p(vo))
    return ko, vo


def get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds) -> np.ndarray:
    keys = np.ascontiguousarray(isect_ids_sorted, dtype=np.int64)
    bins = np.zeros((tile_bounds[0] * tile_bounds[1], 2), np.int32)
    lib().orc_get_tile_bin_edges(C.c_int64(num_intersects), _p(keys), _p(bins))
    return bins


def bin_and_sort_gaussians(num_points, num_intersects, centers, extents, depths, cum_tiles_hit, tile_bounds,
                           block_width, wrapped=False):
    """utils.py:106-162"""
    isect, gids = map_gaussian_to_intersects(num_points, num_intersects, centers, extents, depths, cum_tiles_hit,
                                             tile_bounds, block_width, wrapped)
    isect_s, gids_s = sort_pairs(isect, gids)
    bins = get_tile_bin_edges(num_intersects, isect_s, tile_bounds)
    return isect, gids, isect_s, gids_s, bins


# ----------------------------------------------------------------------------------------------
def sh_num_bases(degree: int) -> int:
    return {0: 1, 1: 4, 2: 9, 3: 16}.get(degree, 25)


def sh_forward(degree, degrees_to_use, viewdirs, coeffs) -> np.ndarray:
    viewdirs, coeffs = _f(viewdirs), _f(coeffs)
    n = coeffs.shape[0]
    out = np.empty((n, 3), np.float32)
    lib().orc_sh_forward(C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use), _p(viewdirs), _p(coeffs), _p(out))
    return out


def sh_backward(degree, degrees_to_use, viewdirs, v_colors) -> np.ndarray:
    viewdirs, v_colors = _f(viewdirs), _f(v_colors)
    n = v_colors.shape[0]
    out = np.zeros((n, sh_num_bases(degree), 3), np.float32)
    lib().orc_sh_backward(C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use), _p(viewdirs), _p(v_colors), _p(out))
    return out


def texture_sample_forward(texture_dims, uvs, texture) -> np.ndarray:
    dims, uvs, texture = _i(texture_dims), _f(uvs), _f(texture)
    nq, ch = uvs.shape[0], texture.shape[1]
    out = np.zeros((nq, ch), np.float32)
    lib().orc_texture_sample_forward(C.c_int(nq), C.c_int(ch), _p(dims), _p(uvs), _p(texture), _p(out))
    return out


def texture_sample_backward(texture_dims, uvs, texture, v_output) -> np.ndarray:
    dims, uvs, texture, v_output = _i(texture_dims), _f(uvs), _f(texture), _f(v_output)
    nq, ch = uvs.shape[0], texture.shape[1]
    out = np.zeros_like(texture)
    lib().orc_texture_sample_backward(C.c_int(nq), C.c_int(ch), _p(dims), _p(uvs), _p(v_output), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
FWD_KEYS = ("out_img", "out_depth", "out_reg", "out_texture", "out_normal", "final_Ts", "final_idx", "depth_idx",
            "out_reg_s")
BWD_KEYS = ("v_colors", "v_opacity", "v_means", "v_scales", "v_quats", "v_uv0", "v_umap", "v_vmap", "v_texture")


def texture_forward(img_height, img_width, block_width, texture_dims, gaussian_ids_sorted, tile_bins, colors,
                    opacities, means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx,
                    cy, settings, background) -> Dict[str, np.ndarray]:
    """texture_forward_tensor, texture.cu:766-901 (same 9 outputs, as a dict)."""
    H, W = int(img_height), int(img_width)
    texture = _f(texture)
    ch = texture.shape[1]
    assert ch <= 64
    o = dict(
        out_img=np.zeros((H, W, 3), np.float32), out_depth=np.zeros((H, W), np.float32),
        out_reg=np.zeros((H, W), np.float32), out_texture=np.zeros((H, W, ch), np.float32),
        out_normal=np.zeros((H, W, 3), np.float32), final_Ts=np.zeros((H, W), np.float32),
        final_idx=np.zeros((H, W), np.int32), depth_idx=np.zeros((H, W), np.int32),
        out_reg_s=np.zeros((H, W, 3), np.float32),
    )
    a = [_i(texture_dims), _i(gaussian_ids_sorted), _i(tile_bins), _f(colors), _f(opacities), _f(means), _f(scales)]
    b = [_f(quats), _f(uv0), _f(umap), _f(vmap), texture, _f(viewmat), _f(c2w)]
    bg = _f(background)
    lib().orc_texture_forward(
        C.c_int(W), C.c_int(H), C.c_int(block_width), C.c_int(ch), *[_p(x) for x in a], C.c_float(glob_scale),
        *[_p(x) for x in b], C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(settings), _p(bg),
        _p(o["out_img"]), _p(o["out_depth"]), _p(o["out_reg"]), _p(o["out_texture"]), _p(o["out_normal"]),
        _p(o["final_Ts"]), _p(o["final_idx"]), _p(o["depth_idx"]), _p(o["out_reg_s"]))
    return o


def texture_backward(img_height, img_width, block_width, texture_dims, gaussian_ids_sorted, tile_bins, colors,
                     opacities, means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx,
                     cy, settings, background, final_Ts, final_idx, depth_idx, final_s, v_out_img, v_out_depth,
                     v_out_reg, v_out_alpha, v_out_texture, v_out_normal) -> Dict[str, np.ndarray]:
    """texture_backward_tensor, texture.cu:915-1053 (same 9 outputs, as a dict)."""
    H, W = int(img_height), int(img_width)
    texture, means = _f(texture), _f(means)
    n, ch = means.shape[0], texture.shape[1]
    o = dict(
        v_colors=np.zeros((n, 3), np.float32), v_opacity=np.zeros((n, 1), np.float32),
        v_means=np.zeros((n, 3), np.float32), v_scales=np.zeros((n, 3), np.float32),
        v_quats=np.zeros((n, 4), np.float32), v_uv0=np.zeros((n, 1, 2), np.float32),
        v_umap=np.zeros((n, 1, 3), np.float32), v_vmap=np.zeros((n, 1, 3), np.float32),
        v_texture=np.zeros_like(texture),
    )
    a = [_i(texture_dims), _i(gaussian_ids_sorted), _i(tile_bins), _f(colors), _f(opacities), means, _f(scales)]
    b = [_f(quats), _f(uv0), _f(umap), _f(vmap), texture, _f(viewmat), _f(c2w)]
    c = [_f(background), _f(final_Ts), _i(final_idx), _i(depth_idx), _f(final_s), _f(v_out_img), _f(v_out_depth),
         _f(v_out_reg), _f(v_out_alpha), _f(v_out_texture), _f(v_out_normal)]
    lib().orc_texture_backward(
        C.c_int(W), C.c_int(H), C.c_int(block_width), C.c_int(ch), *[_p(x) for x in a], C.c_float(glob_scale),
        *[_p(x) for x in b], C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(settings),
        *[_p(x) for x in c], *[_p(o[k]) for k in BWD_KEYS])
    return o


def texture_edit(img_height, img_width, block_width, channels, num_texels, texture_dims, updated_img, updated_alpha,
                 depth_lower, depth_upper, gaussian_ids_sorted, tile_bins, opacities, means, scales, glob_scale, quats,
                 uv0, umap, vmap, viewmat, c2w, fx, fy, cx, cy, settings) -> np.ndarray:
    """texture_edit_tensor, texture_edit.cu:238-354: (num_texels, channels) zero-initialised, 5 channels splatted."""
    H, W = int(img_height), int(img_width)
    assert channels >= 5
    out = np.zeros((int(num_texels), int(channels)), np.float32)
    a = [_i(texture_dims), _f(updated_img), _f(updated_alpha), _f(depth_lower), _f(depth_upper),
         _i(gaussian_ids_sorted), _i(tile_bins), _f(opacities), _f(means), _f(scales)]
    b = [_f(quats), _f(uv0), _f(umap), _f(vmap), _f(viewmat), _f(c2w)]
    lib().orc_texture_edit(
        C.c_int(W), C.c_int(H), C.c_int(block_width), C.c_int(channels), *[_p(x) for x in a], C.c_float(glob_scale),
        *[_p(x) for x in b], C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_int(settings), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# training-step glue
PRE_KEYS = ("scales", "quats", "uv0", "umap", "vmap", "colors", "opacities")
PRE_GRAD_KEYS = ("v_raw_scales", "v_raw_quats", "v_mapping", "v_raw_rgbs", "v_raw_opacities")


def preprocess_forward(raw_scales, raw_quats, mapping, raw_rgbs, raw_opacities) -> Dict[str, np.ndarray]:
    """example.py:126-143 + :162-163; raw_rgbs may be None (colours from SH)."""
    raw_scales, raw_quats, mapping, raw_opacities = _f(raw_scales), _f(raw_quats), _f(mapping), _f(raw_opacities)
    n = raw_scales.shape[0]
    o = dict(scales=np.zeros((n, 3), np.float32), quats=np.zeros((n, 4), np.float32), uv0=np.zeros((n, 1, 2), np.float32),
             umap=np.zeros((n, 1, 3), np.float32), vmap=np.zeros((n, 1, 3), np.float32),
             colors=None if raw_rgbs is None else np.zeros((n, 3), np.float32), opacities=np.zeros((n, 1), np.float32))
    rgb = None if raw_rgbs is None else _f(raw_rgbs)
    lib().orc_preprocess_forward(C.c_int(n), _p(raw_scales), _p(raw_quats), _p(mapping),
                                 None if rgb is None else _p(rgb), _p(raw_opacities), _p(o["scales"]), _p(o["quats"]),
                                 _p(o["uv0"]), _p(o["umap"]), _p(o["vmap"]),
                                 None if rgb is None else _p(o["colors"]), _p(o["opacities"]))
    return o


def preprocess_backward(raw_scales, raw_quats, mapping, raw_rgbs, raw_opacities, v_scales, v_quats, v_uv0, v_umap,
                        v_vmap, v_colors, v_opacity) -> Dict[str, np.ndarray]:
    raw_scales, raw_quats, mapping, raw_opacities = _f(raw_scales), _f(raw_quats), _f(mapping), _f(raw_opacities)
    n = raw_scales.shape[0]
    rgb = None if raw_rgbs is None else _f(raw_rgbs)
    vc = None if raw_rgbs is None else _f(v_colors)
    o = dict(v_raw_scales=np.zeros((n, 3), np.float32), v_raw_quats=np.zeros((n, 4), np.float32),
             v_mapping=np.zeros((n, 1, 4), np.float32), v_raw_rgbs=None if rgb is None else np.zeros((n, 3), np.float32),
             v_raw_opacities=np.zeros((n, 1), np.float32))
    g = [_f(v_scales), _f(v_quats), _f(v_uv0), _f(v_umap), _f(v_vmap)]
    vo = _f(v_opacity)
    lib().orc_preprocess_backward(C.c_int(n), _p(raw_scales), _p(raw_quats), _p(mapping),
                                  None if rgb is None else _p(rgb), _p(raw_opacities), *[_p(x) for x in g],
                                  None if vc is None else _p(vc), _p(vo), _p(o["v_raw_scales"]), _p(o["v_raw_quats"]),
                                  _p(o["v_mapping"]), None if rgb is None else _p(o["v_raw_rgbs"]),
                                  _p(o["v_raw_opacities"]))
    return o


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """torch.optim.Adam update, in place on float32 numpy arrays (step is 1-based)."""
    for a in (params, exp_avg, exp_avg_sq):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    g = _f(grads)
    lib().orc_adam_step(C.c_int64(params.size), _p(params), _p(g), _p(exp_avg), _p(exp_avg_sq), C.c_double(lr),
                        C.c_double(beta1), C.c_double(beta2), C.c_double(eps), C.c_int(step), C.c_float(grad_scale))


# ----------------------------------------------------------------------------------------------
def bin_view(means, scales, glob_scale, quats, viewmat, intrins, img_height, img_width, block_width):
    """project -> aabb -> tile count -> cumsum -> key emit -> sort -> tile ranges, as example.py:146-152
    followed by texture.py:195-222.  Returns a dict with every intermediate."""
    tile_bounds = ((img_width + block_width - 1) // block_width, (img_height + block_width - 1) // block_width, 1)
    _, depths = project_points(means, viewmat, intrins)
    centers, extents = get_aabb_2d(means, scales, glob_scale, quats, viewmat, intrins)
    nth = get_num_tiles_hit_2d(centers, extents, img_height, img_width, block_width)
    m, cum = compute_cumulative_intersects(nth)
    isect, gids, isect_s, gids_s, bins = bin_and_sort_gaussians(
        means.shape[0], m, centers, extents, depths, cum, tile_bounds, block_width)
    return dict(depths=depths, centers=centers, extents=extents, num_tiles_hit=nth, cum_tiles_hit=cum,
                num_intersects=m, isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=isect_s,
                gaussian_ids_sorted=gids_s, tile_bins=bins, tile_bounds=tile_bounds)
