"""Texture editing: splat an edited canvas back into the per-Gaussian textures (mirror of
``gstex_cuda/texture_edit.py``, victor-rong/GStex_cuda).

``texture_edit`` has the reference's signature (texture_edit.py:14-44) and returns ``updated_texture`` of shape
``(sum h*w, texture_info[2])`` whose first five channels are (r*a, g*a, b*a, a, total bilinear weight).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import cuda as _C
from .utils import bin_tiles, compute_cumulative_intersects


def texture_edit(texture_info: Tuple[int, int, int], texture_dims: Tensor, updated_img: Tensor, updated_alpha: Tensor,
                 depth_lower: Tensor, depth_upper: Tensor, centers: Tensor, extents: Tensor, depths: Tensor,
                 num_tiles_hit: Tensor, opacity: Tensor, means: Tensor, scales: Tensor, glob_scale, quats: Tensor,
                 uv0: Tensor, umap: Tensor, vmap: Tensor, viewmat: Tensor, c2w: Tensor, fx: float, fy: float,
                 cx: float, cy: float, img_height: int, img_width: int, block_width: int, settings: int,
                 background: Optional[Tensor] = None, use_torch_impl: bool = False) -> Tensor:
    """Arguments as in the reference (texture_edit.py:46-109).  Not differentiable.  ``settings`` bit 0 = blur,
    bit 1 = ndc (texture_edit.cu:46-47); ``background`` and ``use_torch_impl`` are unused, as upstream."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    num_points = centers.size(0)
    tile_bounds = ((img_width + block_width - 1) // block_width, (img_height + block_width - 1) // block_width, 1)
    block = (block_width, block_width, 1)
    img_size = (img_width, img_height, 1)
    texture_dims = texture_dims.contiguous()
    # texture_edit.py:190 upstream: a blocking .item(), like the cumulative intersection count below
    texture_total_size = int(torch.sum(texture_dims[:, 0] * texture_dims[:, 1]).item())
    num_intersects, _ = compute_cumulative_intersects(num_tiles_hit.contiguous())
    assert num_intersects >= 1  # upstream: `assert False` (texture_edit.py:194-195)
    gaussian_ids_sorted, tile_bins, _, _ = bin_tiles(centers.contiguous(), extents.contiguous(), depths.contiguous(),
                                                     tile_bounds, block_width, num_intersects)
    del num_points
    return _C.texture_edit(tile_bounds, block, img_size, texture_info, texture_total_size, texture_dims,
                           updated_img.contiguous(), updated_alpha.contiguous(), depth_lower.contiguous(),
                           depth_upper.contiguous(), gaussian_ids_sorted, tile_bins, opacity.contiguous(),
                           means.contiguous(), scales.contiguous(), glob_scale, quats.contiguous(), uv0.contiguous(),
                           umap.contiguous(), vmap.contiguous(), viewmat.contiguous(), c2w.contiguous(), fx, fy, cx, cy,
                           settings, None if background is None else background.contiguous())
