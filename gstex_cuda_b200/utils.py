"""Tile binning glue: cumulative intersections, key emission, sort, tile ranges.

Mirror of ``gstex_cuda/utils.py`` (victor-rong/GStex_cuda): same function names, arguments and returns.
``torch.cumsum`` / ``torch.sort`` / ``torch.gather`` of the reference (utils.py:57,159,160) are replaced
by the hand-written scan and stable radix sort of libgstex_b200 (csrc/binning.cu).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from . import cuda as _C


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def get_tile_bin_edges(num_intersects: int, isect_ids_sorted: Tensor, tile_bounds: Tuple[int, int, int]) -> Tensor:
    """utils.py:11-37: (start, end) range of sorted intersections per tile; not differentiable."""
    return _C.get_tile_bin_edges(num_intersects, isect_ids_sorted.contiguous(), tile_bounds)


def cumsum_i32(values: Tensor) -> Tensor:
    """Inclusive int32 prefix sum on the GPU (csrc/binning.cu scan kernels)."""
    if not values.is_cuda:
        raise RuntimeError("num_tiles_hit must be a CUDA tensor")
    v = values.contiguous()
    if v.dtype != torch.int32:
        v = v.to(torch.int32)
    n = v.numel()
    out = torch.empty((n,), dtype=torch.int32, device=v.device)
    if n == 0:
        return out
    lib = _lib.load()
    temp = torch.empty((lib.gstex_scan_temp_bytes(n),), dtype=torch.uint8, device=v.device)
    with torch.cuda.device(v.device):
        rc = lib.gstex_cumsum_i32(n, v.data_ptr(), out.data_ptr(), temp.data_ptr(), temp.numel(), _stream(v.device))
    _lib.check(rc, "cumsum_i32")
    return out


def compute_cumulative_intersects(num_tiles_hit: Tensor) -> Tuple[int, Tensor]:
    """utils.py:40-59: (num_intersects, inclusive cumsum).  The ``.item()`` is the reference's one
    device->host sync per iteration; the fused path (gstex_cuda_b200.pipeline) avoids it."""
    cum_tiles_hit = cumsum_i32(num_tiles_hit.reshape(-1))
    num_intersects = int(cum_tiles_hit[-1].item()) if cum_tiles_hit.numel() else 0
    return num_intersects, cum_tiles_hit


def map_gaussian_to_intersects(num_points: int, num_intersects: int, centers: Tensor, extents: Tensor,
                               depths: Tensor, cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int],
                               block_size: int, wrapped: bool = False) -> Tuple[Tensor, Tensor]:
    """utils.py:61-104: (tile | depth) int64 keys and Gaussian ids, unsorted."""
    return _C.map_gaussian_to_intersects(num_points, num_intersects, centers.contiguous(), extents.contiguous(),
                                         depths.contiguous(), cum_tiles_hit.contiguous(), tile_bounds, block_size,
                                         wrapped)


def sort_pairs(keys: Tensor, values: Tensor, end_bit: int = 64) -> Tuple[Tensor, Tensor]:
    """Stable ascending sort of int64 keys carrying int32 values (replaces torch.sort + torch.gather)."""
    if not keys.is_cuda or keys.dtype != torch.int64 or values.dtype != torch.int32:
        raise RuntimeError("sort_pairs expects CUDA int64 keys and int32 values")
    keys, values = keys.contiguous(), values.contiguous()
    m = keys.numel()
    keys_out, vals_out = torch.empty_like(keys), torch.empty_like(values)
    if m == 0:
        return keys_out, vals_out
    lib = _lib.load()
    temp = torch.empty((lib.gstex_sort_temp_bytes(m),), dtype=torch.uint8, device=keys.device)
    with torch.cuda.device(keys.device):
        rc = lib.gstex_sort_pairs(m, keys.data_ptr(), values.data_ptr(), keys_out.data_ptr(), vals_out.data_ptr(),
                                  int(end_bit), 0, temp.data_ptr(), temp.numel(), _stream(keys.device))
    _lib.check(rc, "sort_pairs")
    return keys_out, vals_out


def bin_and_sort_gaussians(num_points: int, num_intersects: int, centers: Tensor, extents: Tensor, depths: Tensor,
                           cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_size: int,
                           wrapped: bool = False):
    """utils.py:106-162: returns (isect_ids_unsorted, gaussian_ids_unsorted, isect_ids_sorted,
    gaussian_ids_sorted, tile_bins)."""
    isect_ids, gaussian_ids = map_gaussian_to_intersects(num_points, num_intersects, centers, extents, depths,
                                                         cum_tiles_hit, tile_bounds, block_size, wrapped=wrapped)
    isect_ids_sorted, gaussian_ids_sorted = sort_pairs(isect_ids, gaussian_ids)
    tile_bins = get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds)
    return isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins


def bin_tiles(centers: Tensor, extents: Tensor, depths: Tensor, tile_bounds: Tuple[int, int, int], block_size: int,
              capacity: int, want_isect_ids: bool = False):
    """Fused tile binning (csrc/binning_tiles.cu): bucket by tile, then sort each tile's list in shared memory.

    Returns ``(gaussian_ids_sorted[capacity], tile_bins, num_intersects_dev[1], isect_ids_sorted or None)`` -
    bit-identical to ``bin_and_sort_gaussians`` (utils.py:106-162 upstream) on the first ``num_intersects`` entries,
    without the cumulative sum, the 64-bit global sort or any host synchronisation."""
    if not (centers.is_cuda and extents.is_cuda and depths.is_cuda):
        raise RuntimeError("bin_tiles expects CUDA tensors")
    centers, extents, depths = centers.contiguous(), extents.contiguous(), depths.contiguous()
    dev = centers.device
    n = centers.shape[0]
    tx, ty = int(tile_bounds[0]), int(tile_bounds[1])
    lib = _lib.load()
    ids = torch.empty((capacity,), dtype=torch.int32, device=dev)
    isect = torch.empty((capacity,), dtype=torch.int64, device=dev) if want_isect_ids else None
    bins = torch.empty((tx * ty, 2), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    temp = torch.empty((lib.gstex_bin_tiles_temp_bytes(tx * ty, capacity),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.gstex_bin_tiles(n, centers.data_ptr(), extents.data_ptr(), depths.data_ptr(), tx, ty, int(block_size),
                                 int(capacity), ids.data_ptr(), isect.data_ptr() if isect is not None else 0,
                                 bins.data_ptr(), count.data_ptr(), 0, temp.data_ptr(), temp.numel(), _stream(dev))
    _lib.check(rc, "bin_tiles")
    return ids, bins, count, isect
