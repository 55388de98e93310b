"""Backend module: the ten callables of the reference's pybind module, on top of the C ABI.

Mirror of ``gstex_cuda/cuda/__init__.py:14-27`` + ``csrc/ext.cpp:8-25`` of victor-rong/GStex_cuda: same
names, same positional arguments, same return tuples / dtypes / shapes, zero-filled where the reference
zero-fills (``torch::zeros``).  Inputs must be CUDA tensors and contiguous, as ``CHECK_INPUT`` demands
(``csrc/bindings.h:10-15``); violations raise ``RuntimeError`` like the reference's ``TORCH_CHECK``.

Everything runs on the current CUDA stream of the tensors' device (the reference launches on the legacy
default stream and guards the device only in 4 of its 10 wrappers).
"""
from __future__ import annotations

from typing import Tuple

import torch

from .. import _lib

__all__ = [
    "compute_sh_forward", "compute_sh_backward", "map_gaussian_to_intersects", "get_tile_bin_edges", "get_aabb_2d",
    "texture_forward", "texture_backward", "texture_sample_forward", "texture_sample_backward", "texture_edit",
]


def _chk(name: str, t: torch.Tensor, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected scalar type {dtype} for {name} but found {t.dtype}")
    return t


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def num_sh_bases(degree: int) -> int:
    return {0: 1, 1: 4, 2: 9, 3: 16}.get(int(degree), 25)


# ------------------------------------------------------------------------------------------------
def compute_sh_forward(num_points, degree, degrees_to_use, viewdirs, coeffs):
    """bindings.cu:18-43"""
    _chk("viewdirs", viewdirs, torch.float32), _chk("coeffs", coeffs, torch.float32)
    nb = num_sh_bases(degree)
    if coeffs.dim() != 3 or coeffs.shape[0] != num_points or coeffs.shape[1] != nb or coeffs.shape[2] != 3:
        raise RuntimeError("coeffs must have dimensions (N, D, 3)")
    colors = torch.empty((num_points, 3), dtype=torch.float32, device=coeffs.device)
    with torch.cuda.device(coeffs.device):
        rc = _lib.load().gstex_sh_forward(int(num_points), int(degree), int(degrees_to_use), _p(viewdirs), _p(coeffs),
                                          _p(colors), _stream(coeffs.device))
    _lib.check(rc, "compute_sh_forward")
    return colors


def compute_sh_backward(num_points, degree, degrees_to_use, viewdirs, v_colors):
    """bindings.cu:45-75"""
    _chk("viewdirs", viewdirs, torch.float32), _chk("v_colors", v_colors, torch.float32)
    if viewdirs.dim() != 2 or viewdirs.shape[0] != num_points or viewdirs.shape[1] != 3:
        raise RuntimeError("viewdirs must have dimensions (N, 3)")
    if v_colors.dim() != 2 or v_colors.shape[0] != num_points or v_colors.shape[1] != 3:
        raise RuntimeError("v_colors must have dimensions (N, 3)")
    v_coeffs = torch.empty((num_points, num_sh_bases(degree), 3), dtype=torch.float32, device=v_colors.device)
    with torch.cuda.device(v_colors.device):
        rc = _lib.load().gstex_sh_backward(int(num_points), int(degree), int(degrees_to_use), _p(viewdirs),
                                           _p(v_colors), _p(v_coeffs), 0, _stream(v_colors.device))
    _lib.check(rc, "compute_sh_backward")
    return v_coeffs


def map_gaussian_to_intersects(num_points, num_intersects, centers, extents, depths, cum_tiles_hit, tile_bounds,
                               block_width, wrapped=False) -> Tuple[torch.Tensor, torch.Tensor]:
    """bindings.cu:77-121.  ``wrapped`` selects the torus tile boxes of forward.cu:34-36, 53-62."""
    _chk("centers", centers, torch.float32), _chk("extents", extents, torch.float32)
    _chk("depths", depths, torch.float32), _chk("cum_tiles_hit", cum_tiles_hit, torch.int32)
    dev = centers.device
    isect = torch.zeros((num_intersects,), dtype=torch.int64, device=dev)
    gids = torch.zeros((num_intersects,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        lib = _lib.load()
        fn = lib.gstex_map_gaussian_to_intersects_wrapped if wrapped else lib.gstex_map_gaussian_to_intersects
        rc = fn(
            int(num_points), int(num_intersects), _p(centers), _p(extents), _p(depths), _p(cum_tiles_hit),
            int(tile_bounds[0]), int(tile_bounds[1]), int(block_width), _p(isect), _p(gids), _stream(dev))
    _lib.check(rc, "map_gaussian_to_intersects")
    return isect, gids


def get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds) -> torch.Tensor:
    """bindings.cu:123-140"""
    _chk("isect_ids_sorted", isect_ids_sorted, torch.int64)
    dev = isect_ids_sorted.device
    bins = torch.zeros((int(tile_bounds[0]) * int(tile_bounds[1]), 2), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().gstex_get_tile_bin_edges(int(num_intersects), _p(isect_ids_sorted), _p(bins), 0, _stream(dev))
    _lib.check(rc, "get_tile_bin_edges")
    return bins


def get_aabb_2d(means, scales, glob_scale, quats, viewmat, fx, fy, cx, cy):
    """get_aabb_2d.cu:91-125"""
    for n, t in (("means", means), ("scales", scales), ("quats", quats), ("viewmat", viewmat)):
        _chk(n, t, torch.float32)
    n = means.shape[0]
    centers = torch.zeros((n, 2), dtype=torch.float32, device=means.device)
    extents = torch.zeros((n, 2), dtype=torch.float32, device=means.device)
    with torch.cuda.device(means.device):
        rc = _lib.load().gstex_get_aabb_2d(n, _p(means), _p(scales), float(glob_scale), _p(quats), _p(viewmat), float(fx),
                                           float(fy), float(cx), float(cy), _p(centers), _p(extents),
                                           _stream(means.device))
    _lib.check(rc, "get_aabb_2d")
    return centers, extents


# ------------------------------------------------------------------------------------------------
_RASTER_IN = ("texture_dims", "gaussian_ids_sorted", "tile_bins", "colors", "opacities", "means", "scales", "quats",
              "uv0", "umap", "vmap", "texture", "viewmat", "c2w", "background")


def _check_raster_inputs(texture_dims, gaussian_ids_sorted, tile_bins, colors, opacities, means, scales, quats, uv0,
                         umap, vmap, texture, viewmat, c2w, background):
    _chk("texture_dims", texture_dims, torch.int32)
    if gaussian_ids_sorted is not None:  # None: the caller bins after this check (texture.py)
        _chk("gaussian_ids_sorted", gaussian_ids_sorted, torch.int32)
        _chk("tile_bins", tile_bins, torch.int32)
    for n, t in (("colors", colors), ("opacities", opacities), ("means", means), ("scales", scales), ("quats", quats),
                 ("uv0", uv0), ("umap", umap), ("vmap", vmap), ("texture", texture), ("viewmat", viewmat),
                 ("c2w", c2w), ("background", background)):
        _chk(n, t, torch.float32)
    if colors.dim() != 2 or colors.shape[1] != 3:
        raise RuntimeError("colors must have dimensions (N, 3)")  # float3 casts upstream, texture.cu:871
    if background.numel() < 3:
        raise RuntimeError("background must hold at least 3 floats")


def _forward_scratch(n, num_texels, channels, num_intersects, dev) -> torch.Tensor:
    nbytes = _lib.load().gstex_texture_forward_temp_bytes(n, num_texels, channels, num_intersects)
    return torch.empty((nbytes,), dtype=torch.uint8, device=dev)


def _texture_forward(keep_scratch, tile_bounds, block, img_size, texture_info, texture_dims, gaussian_ids_sorted,
                     tile_bins, colors, opacities, means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat,
                     c2w, fx, fy, cx, cy, settings, background):
    _check_raster_inputs(texture_dims, gaussian_ids_sorted, tile_bins, colors, opacities, means, scales, quats, uv0,
                         umap, vmap, texture, viewmat, c2w, background)
    dev = means.device
    W, H = int(img_size[0]), int(img_size[1])
    bw = int(block[0])
    n, X, C = means.shape[0], texture.shape[0], int(texture_info[2])
    if texture.dim() != 2 or texture.shape[1] != C:
        raise RuntimeError(f"texture must have dimensions (X, {C})")
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    out_img, out_depth, out_reg = torch.empty((H, W, 3), **f32), torch.empty((H, W), **f32), torch.empty((H, W), **f32)
    out_texture, out_normal = torch.empty((H, W, C), **f32), torch.empty((H, W, 3), **f32)
    final_Ts, final_idx, depth_idx = torch.empty((H, W), **f32), torch.empty((H, W), **i32), torch.empty((H, W), **i32)
    out_reg_s = torch.empty((H, W, 3), **f32)
    # the blend masks (32 B per list entry) are only kept for a caller that will hand the scratch to the backward
    M = gaussian_ids_sorted.shape[0] if keep_scratch else 0
    scratch = _forward_scratch(n, X, C, M, dev)
    with torch.cuda.device(dev):
        rc = _lib.load().gstex_texture_forward(
            H, W, bw, n, X, C, M, _p(texture_dims), _p(gaussian_ids_sorted), _p(tile_bins), _p(colors), _p(opacities),
            _p(means), _p(scales), float(glob_scale), _p(quats), _p(uv0), _p(umap), _p(vmap), _p(texture), _p(viewmat),
            _p(c2w), float(fx), float(fy), float(cx), float(cy), int(settings), _p(background), _p(out_img),
            _p(out_depth), _p(out_reg), _p(out_texture), _p(out_normal), _p(final_Ts), _p(final_idx), _p(depth_idx),
            _p(out_reg_s), _p(scratch), scratch.numel(), _stream(dev))
    _lib.check(rc, "texture_forward")
    return (out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, final_idx, depth_idx, out_reg_s), scratch


def texture_forward_ex(*args):
    """texture_forward plus the forward scratch (packed records, padded texture, blend masks): handing it to
    ``texture_backward(..., _fwd_scratch=scratch)`` saves that call the re-derivation of the three."""
    return _texture_forward(True, *args)


def texture_forward(*args):
    """texture_forward_tensor, texture.cu:766-901: returns the same 9-tuple.  Keeps no state for the backward."""
    return _texture_forward(False, *args)[0]


def texture_backward(img_height, img_width, block_width, texture_info, texture_dims, gaussian_ids_sorted, tile_bins,
                     colors, opacities, means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy,
                     cx, cy, settings, background, final_Ts, final_idx, depth_idx, final_s, v_output, v_output_depth,
                     v_output_reg, v_output_alpha, v_output_texture, v_output_normal, _fwd_scratch=None):
    """texture_backward_tensor, texture.cu:915-1053: returns the same 9-tuple of gradients, a pure function of its
    arguments like upstream.

    ``_fwd_scratch`` (keyword, optional, not in the reference signature) is the scratch tensor returned by
    ``texture_forward_ex`` for the same inputs; with it the call skips re-packing the records / texture and re-deriving
    the blend masks from ``final_Ts`` / ``final_idx`` (a cull + alpha walk over the lists, no second forward pass).
    """
    _check_raster_inputs(texture_dims, gaussian_ids_sorted, tile_bins, colors, opacities, means, scales, quats, uv0,
                         umap, vmap, texture, viewmat, c2w, background)
    _chk("final_Ts", final_Ts, torch.float32), _chk("final_idx", final_idx, torch.int32)
    _chk("depth_idx", depth_idx, torch.int32), _chk("final_s", final_s, torch.float32)
    for n_, t in (("v_output", v_output), ("v_output_depth", v_output_depth), ("v_output_reg", v_output_reg),
                  ("v_output_alpha", v_output_alpha), ("v_output_texture", v_output_texture),
                  ("v_output_normal", v_output_normal)):
        _chk(n_, t, torch.float32)
    dev = means.device
    H, W, bw = int(img_height), int(img_width), int(block_width)
    n, X, C = means.shape[0], texture.shape[0], int(texture_info[2])
    nprob = int(texture_info[1])
    M = gaussian_ids_sorted.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    v_colors, v_opacity = torch.empty((n, 3), **f32), torch.empty((n, 1), **f32)
    v_means, v_scales, v_quats = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty((n, 4), **f32)
    # (n, num_probs, k) zero-filled as upstream (texture.cu:1004-1006).  The reference kernels address these arrays
    # (and uv0 / umap / vmap) flat, as float2 / float3 element g (texture.cu:743-756), whatever num_probs is: Gaussian
    # g's gradient lands at flat element g, and so it does here
    if nprob == 1:
        v_uv0, v_umap, v_vmap = torch.empty((n, 1, 2), **f32), torch.empty((n, 1, 3), **f32), torch.empty((n, 1, 3), **f32)
        o_uv0, o_umap, o_vmap = v_uv0, v_umap, v_vmap
    else:
        v_uv0, v_umap, v_vmap = torch.zeros((n, nprob, 2), **f32), torch.zeros((n, nprob, 3), **f32), torch.zeros((n, nprob, 3), **f32)
        o_uv0, o_umap, o_vmap = torch.empty((n, 2), **f32), torch.empty((n, 3), **f32), torch.empty((n, 3), **f32)
    v_texture = torch.empty((X, C), **f32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        nbytes = (lib.gstex_texture_backward_temp_bytes(n, X, C) if _fwd_scratch is not None
                  else lib.gstex_texture_backward_stateless_temp_bytes(n, X, C, M))
        temp = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        rc = lib.gstex_texture_backward(
            H, W, bw, n, X, C, M, _p(texture_dims), _p(gaussian_ids_sorted), _p(tile_bins), _p(colors), _p(opacities),
            _p(means), _p(scales), float(glob_scale), _p(quats), _p(uv0), _p(umap), _p(vmap), _p(texture), _p(viewmat),
            _p(c2w), float(fx), float(fy), float(cx), float(cy), int(settings), _p(background), _p(final_Ts),
            _p(final_idx), _p(depth_idx), _p(final_s), _p(v_output), _p(v_output_depth), _p(v_output_reg),
            _p(v_output_alpha), _p(v_output_texture), _p(v_output_normal), _p(v_colors), _p(v_opacity), _p(v_means),
            _p(v_scales), _p(v_quats), _p(o_uv0), _p(o_umap), _p(o_vmap), _p(v_texture), 0, _p(_fwd_scratch), _p(temp),
            temp.numel(), _stream(dev))
    _lib.check(rc, "texture_backward")
    if nprob != 1:
        v_uv0.view(-1)[:2 * n] = o_uv0.view(-1)
        v_umap.view(-1)[:3 * n] = o_umap.view(-1)
        v_vmap.view(-1)[:3 * n] = o_vmap.view(-1)
    return v_colors, v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap, v_texture


# ------------------------------------------------------------------------------------------------
def texture_sample_forward(texture_info, texture_dims, uvs, texture):
    """texture_sample.cu:72-105"""
    _chk("texture_dims", texture_dims, torch.int32), _chk("uvs", uvs, torch.float32)
    _chk("texture", texture, torch.float32)
    nq, C = uvs.shape[0], int(texture_info[2])
    out = torch.empty((nq, C), dtype=torch.float32, device=texture.device)
    with torch.cuda.device(texture.device):
        rc = _lib.load().gstex_texture_sample_forward(nq, C, _p(texture_dims), _p(uvs), _p(texture), _p(out),
                                                      _stream(texture.device))
    _lib.check(rc, "texture_sample_forward")
    return out


def texture_sample_backward(texture_info, texture_dims, uvs, texture, v_output):
    """texture_sample.cu:107-141 (implemented as the intended scatter; see include/gstex_b200.h)."""
    _chk("texture_dims", texture_dims, torch.int32), _chk("uvs", uvs, torch.float32)
    _chk("texture", texture, torch.float32), _chk("v_output", v_output, torch.float32)
    nq, C = uvs.shape[0], int(texture_info[2])
    v_texture = torch.zeros((texture.shape[0], C), dtype=torch.float32, device=texture.device)
    with torch.cuda.device(texture.device):
        rc = _lib.load().gstex_texture_sample_backward(nq, C, _p(texture_dims), _p(uvs), _p(v_output), _p(v_texture),
                                                       _stream(texture.device))
    _lib.check(rc, "texture_sample_backward")
    return v_texture


def texture_edit(tile_bounds, block, img_size, texture_info, texture_total_size, texture_dims, updated_img,
                 updated_alpha, depth_lower, depth_upper, gaussian_ids_sorted, tile_bins, opacities, means, scales,
                 glob_scale, quats, uv0, umap, vmap, viewmat, c2w, fx, fy, cx, cy, settings, background):
    """texture_edit_tensor, texture_edit.cu:238-354: returns updated_texture (texture_total_size, texture_info[2]),
    zero-initialised, channels 0-4 = (r*a, g*a, b*a, a, weight) splatted with the bilinear texel weights.
    ``settings``: bit 0 = blur, bit 1 = ndc (texture_edit.cu:46-47).  ``background`` is unused, as upstream."""
    _chk("texture_dims", texture_dims, torch.int32)
    _chk("gaussian_ids_sorted", gaussian_ids_sorted, torch.int32), _chk("tile_bins", tile_bins, torch.int32)
    for n_, t in (("updated_img", updated_img), ("updated_alpha", updated_alpha), ("depth_lower", depth_lower),
                  ("depth_upper", depth_upper), ("opacities", opacities), ("means", means), ("scales", scales),
                  ("quats", quats), ("uv0", uv0), ("umap", umap), ("vmap", vmap), ("viewmat", viewmat), ("c2w", c2w)):
        _chk(n_, t, torch.float32)
    if background is not None:
        _chk("background", background)
    dev = means.device
    W, H, bw = int(img_size[0]), int(img_size[1]), int(block[0])
    if updated_img.numel() != H * W * 3 or updated_alpha.numel() != H * W:
        raise RuntimeError("updated_img must be (H, W, 3) and updated_alpha (H, W, 1)")
    if depth_lower.numel() != H * W or depth_upper.numel() != H * W:
        raise RuntimeError("depth_lower / depth_upper must be (H, W)")
    n, X, C = means.shape[0], int(texture_total_size), int(texture_info[2])
    out = torch.empty((X, C), dtype=torch.float32, device=dev)
    lib = _lib.load()
    temp = torch.empty((lib.gstex_texture_edit_temp_bytes(n),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.gstex_texture_edit(
            H, W, bw, n, X, C, _p(texture_dims), _p(updated_img), _p(updated_alpha), _p(depth_lower), _p(depth_upper),
            _p(gaussian_ids_sorted), _p(tile_bins), _p(opacities), _p(means), _p(scales), float(glob_scale), _p(quats),
            _p(uv0), _p(umap), _p(vmap), _p(viewmat), _p(c2w), float(fx), float(fy), float(cx), float(cy),
            int(settings), _p(out), _p(temp), temp.numel(), _stream(dev))
    _lib.check(rc, "texture_edit")
    return out
