"""Differentiable textured-2DGS rasterisation (mirror of ``gstex_cuda/texture.py``, victor-rong/GStex_cuda).

``texture_gaussians`` has the reference's signature and return tuple
``(out_img, out_depth, out_reg, out_alpha, out_texture, out_normal)`` (texture.py:14-150, :286-289).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import cuda as _C
from .utils import bin_and_sort_gaussians, bin_tiles, compute_cumulative_intersects  # noqa: F401


def texture_gaussians(texture_info: Tuple[int, int, int], texture_dims: Tensor, centers: Tensor, extents: Tensor,
                      depths: Tensor, num_tiles_hit: Tensor, colors: Tensor, opacity: Tensor, means: Tensor,
                      scales: Tensor, glob_scale, quats: Tensor, uv0: Tensor, umap: Tensor, vmap: Tensor,
                      texture: Tensor, viewmat: Tensor, c2w: Tensor, fx: float, fy: float, cx: float, cy: float,
                      img_height: int, img_width: int, block_width: int, settings: int,
                      background: Optional[Tensor] = None, use_torch_impl: bool = False):
    """Rasterise textured 2D Gaussians; differentiable w.r.t. colors, opacity, means, scales, quats, uv0,
    umap, vmap and texture.  Arguments, defaults and outputs as in the reference (texture.py:14-150)."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    if colors.dtype == torch.uint8:
        colors = colors.float() / 255
    if background is not None:
        assert background.shape[0] == colors.shape[-1], (
            f"incorrect shape of background color tensor, expected shape {colors.shape[-1]}")
    else:
        background = torch.ones(colors.shape[-1], dtype=torch.float32, device=colors.device)
    if colors.ndimension() != 2:
        raise ValueError("colors must have dimensions (N, D)")
    if use_torch_impl:
        raise NotImplementedError("the pure-PyTorch rasteriser of the reference (gstex_cuda/_torch_impl.py) is its "
                                  "CPU twin, not part of this build; use the CUDA path")
    return _TextureGaussians.apply(
        texture_info, texture_dims.contiguous(), centers.contiguous(), extents.contiguous(), depths.contiguous(),
        num_tiles_hit.contiguous(), colors.contiguous(), opacity.contiguous(), means.contiguous(), scales.contiguous(),
        glob_scale, quats.contiguous(), uv0.contiguous(), umap.contiguous(), vmap.contiguous(), texture.contiguous(),
        viewmat.contiguous(), c2w.contiguous(), fx, fy, cx, cy, img_height, img_width, block_width, settings,
        background.contiguous())


class _TextureGaussians(Function):
    @staticmethod
    def forward(ctx, texture_info, texture_dims, centers, extents, depths, num_tiles_hit, colors, opacity, means,
                scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy, img_height,
                img_width, block_width, settings, background):
        num_points = centers.size(0)
        tile_bounds = ((img_width + block_width - 1) // block_width, (img_height + block_width - 1) // block_width, 1)
        block = (block_width, block_width, 1)
        img_size = (img_width, img_height, 1)
        dev = centers.device
        num_intersects, cum_tiles_hit = compute_cumulative_intersects(num_tiles_hit)
        ctx.num_intersects = num_intersects
        C = int(texture_info[2])
        if num_intersects < 1:
            # upstream leaves several outputs undefined in this branch (texture.py:197-205, :254-289);
            # we return the background-only image and zeros
            f32 = dict(dtype=torch.float32, device=dev)
            out_img = torch.ones(img_height, img_width, colors.shape[-1], **f32) * background
            zeros = torch.zeros(img_height, img_width, **f32)
            ctx.save_for_backward(colors, opacity, means, scales, quats, uv0, umap, vmap, texture)
            return (out_img, zeros, zeros.clone(), zeros.clone(), torch.zeros(img_height, img_width, C, **f32),
                    torch.zeros(img_height, img_width, 3, **f32))
        # same gaussian_ids_sorted / tile_bins as bin_and_sort_gaussians (utils.py:106-162 upstream), from the fused
        # bucket-by-tile + per-tile sort (csrc/binning_tiles.cu) instead of the global 64-bit key sort
        gaussian_ids_sorted, tile_bins, _, _ = bin_tiles(centers, extents, depths, tile_bounds, block_width,
                                                         num_intersects)
        outputs, scratch = _C.texture_forward_ex(
            tile_bounds, block, img_size, texture_info, texture_dims, gaussian_ids_sorted, tile_bins, colors, opacity,
            means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy, settings,
            background)
        out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, final_idx, depth_idx, out_reg_s = outputs
        ctx.img_width, ctx.img_height, ctx.block_width = img_width, img_height, block_width
        ctx.texture_info, ctx.settings, ctx.glob_scale = texture_info, settings, glob_scale
        ctx.intr = (fx, fy, cx, cy)
        ctx.save_for_backward(texture_dims, gaussian_ids_sorted, tile_bins, colors, opacity, means, scales, quats, uv0,
                              umap, vmap, texture, viewmat, c2w, background, final_Ts, final_idx, depth_idx, out_reg_s,
                              scratch)
        out_alpha = 1 - final_Ts
        return out_img, out_depth, out_reg, out_alpha, out_texture, out_normal

    @staticmethod
    def backward(ctx, v_out_img, v_out_depth, v_out_reg, v_out_alpha, v_out_texture, v_out_normal):
        none18 = [None] * 27
        if ctx.num_intersects < 1:
            colors, opacity, means, scales, quats, uv0, umap, vmap, texture = ctx.saved_tensors
            grads = [torch.zeros_like(t) for t in (colors, opacity, means, scales, quats, uv0, umap, vmap, texture)]
        else:
            (texture_dims, gaussian_ids_sorted, tile_bins, colors, opacity, means, scales, quats, uv0, umap, vmap,
             texture, viewmat, c2w, background, final_Ts, final_idx, depth_idx, out_reg_s, scratch) = ctx.saved_tensors
            H, W = ctx.img_height, ctx.img_width
            dev = means.device

            def dense(v, shape):
                return torch.zeros(shape, dtype=torch.float32, device=dev) if v is None else v.contiguous()

            C = int(ctx.texture_info[2])
            fx, fy, cx, cy = ctx.intr
            grads = _C.texture_backward(
                H, W, ctx.block_width, ctx.texture_info, texture_dims, gaussian_ids_sorted, tile_bins, colors, opacity,
                means, scales, ctx.glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy,
                ctx.settings, background, final_Ts, final_idx, depth_idx, out_reg_s, dense(v_out_img, (H, W, 3)),
                dense(v_out_depth, (H, W)), dense(v_out_reg, (H, W)), dense(v_out_alpha, (H, W)),
                dense(v_out_texture, (H, W, C)), dense(v_out_normal, (H, W, 3)), _fwd_scratch=scratch)
        v_colors, v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap, v_texture = grads
        none18[6], none18[7], none18[8], none18[9] = v_colors, v_opacity, v_means, v_scales
        none18[11], none18[12], none18[13], none18[14], none18[15] = v_quats, v_uv0, v_umap, v_vmap, v_texture
        return tuple(none18)
