"""Differentiable textured-2DGS rasterisation (mirror of ``gstex_cuda/texture.py``, victor-rong/GStex_cuda).

``texture_gaussians`` has the reference's signature and return tuple
``(out_img, out_depth, out_reg, out_alpha, out_texture, out_normal)`` (texture.py:14-150, :286-289).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib
from . import cuda as _C
from .utils import bin_and_sort_gaussians, bin_tiles, compute_cumulative_intersects, cumsum_i32  # noqa: F401


def texture_gaussians(texture_info: Tuple[int, int, int], texture_dims: Tensor, centers: Tensor, extents: Tensor,
                      depths: Tensor, num_tiles_hit: Tensor, colors: Tensor, opacity: Tensor, means: Tensor,
                      scales: Tensor, glob_scale, quats: Tensor, uv0: Tensor, umap: Tensor, vmap: Tensor,
                      texture: Tensor, viewmat: Tensor, c2w: Tensor, fx: float, fy: float, cx: float, cy: float,
                      img_height: int, img_width: int, block_width: int, settings: int,
                      background: Optional[Tensor] = None, use_torch_impl: bool = False,
                      max_intersects: Optional[int] = None, texture_grad: Optional[Tensor] = None):
    """Rasterise textured 2D Gaussians; differentiable w.r.t. colors, opacity, means, scales, quats, uv0,
    umap, vmap and texture.  Arguments, defaults and outputs as in the reference (texture.py:14-150).

    ``max_intersects`` (extension, opt-in): a capacity for the sorted intersection list.  The reference synchronises
    the host once per call to read the intersection count (``cum_tiles_hit[-1].item()``, utils.py:58) because its
    buffers are sized by it; with a capacity the count stays on the device and the call never waits for the GPU.
    Intersections beyond the capacity are dropped: ``last_intersect_count()`` returns the count of the last call.

    ``texture_grad`` (extension, opt-in): a float32 tensor shaped like ``texture`` that the backward pass ADDS the texel
    gradients into, instead of returning them to autograd (``texture.grad`` is then left untouched: pass the buffer the
    optimiser or the data-parallel reduction reads).  A multi-view step otherwise pays, per view, for a zero-filled
    gradient the size of the texture plus autograd's ``grad += new`` pass over it."""
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
    if texture_grad is not None:
        if use_torch_impl:
            raise ValueError("texture_grad is an option of the CUDA path")
        if not (texture_grad.is_cuda and texture_grad.dtype == torch.float32 and texture_grad.is_contiguous()
                and texture_grad.shape == texture.shape and texture_grad.device == texture.device):
            raise ValueError("texture_grad must be a contiguous float32 CUDA tensor shaped like texture")
    if use_torch_impl:
        # the slow pure-PyTorch checker (reference texture.py:411-514 over _torch_impl.texture_forward): same binning,
        # torch autograd instead of the backward kernel
        return _texture_gaussians_torch(texture_info, texture_dims, centers, extents, depths, num_tiles_hit, colors, opacity,
                                        means, scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx,
                                        cy, img_height, img_width, block_width, settings, background)
    return _TextureGaussians.apply(
        texture_info, texture_dims.contiguous(), centers.contiguous(), extents.contiguous(), depths.contiguous(),
        num_tiles_hit.contiguous(), colors.contiguous(), opacity.contiguous(), means.contiguous(), scales.contiguous(),
        glob_scale, quats.contiguous(), uv0.contiguous(), umap.contiguous(), vmap.contiguous(), texture.contiguous(),
        viewmat.contiguous(), c2w.contiguous(), fx, fy, cx, cy, img_height, img_width, block_width, settings,
        background.contiguous(), None if max_intersects is None else int(max_intersects), texture_grad)


_LAST_COUNT = {}


def last_intersect_count(device=None) -> int:
    """Intersection count of the last ``texture_gaussians(..., max_intersects=...)`` call on the device (synchronises).
    Compare it with the capacity passed: a larger count means intersections were dropped."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = _LAST_COUNT.get(dev.index if dev.index is not None else torch.cuda.current_device())
    return -1 if t is None else int(t.item())


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def _texture_gaussians_torch(texture_info, texture_dims, centers, extents, depths, num_tiles_hit, colors, opacity, means,
                             scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy, img_height,
                             img_width, block_width, settings, background):
    """``use_torch_impl=True`` (reference texture.py:411-514): binning through the library (not differentiable, as
    upstream), rasterisation in plain torch ops (``_torch_impl.texture_forward``), gradients by torch autograd."""
    from . import _torch_impl as _T

    H, W, bw = int(img_height), int(img_width), int(block_width)
    tile_bounds = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    with torch.no_grad():
        num_intersects, cum = compute_cumulative_intersects(num_tiles_hit.contiguous())
    if num_intersects < 1:
        img = torch.ones(H, W, colors.shape[-1], device=centers.device) * background
        z = torch.zeros(H, W, device=centers.device)
        return (img, z, z.clone(), z.clone(), torch.zeros(H, W, int(texture_info[2]), device=centers.device),
                torch.zeros(H, W, 3, device=centers.device))
    with torch.no_grad():
        _, _, _, ids_sorted, tile_bins = bin_and_sort_gaussians(means.shape[0], num_intersects, centers.contiguous(),
                                                                extents.contiguous(), depths.contiguous(), cum,
                                                                tile_bounds, bw)
    out_img, out_depth, out_reg, out_texture, out_normal, final_Ts, _ = _T.texture_forward(
        tile_bounds, (bw, bw, 1), (W, H, 1), texture_info, texture_dims, ids_sorted, tile_bins, colors, opacity, means,
        scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy, int(settings), background)
    return out_img, out_depth, out_reg, 1 - final_Ts, out_texture, out_normal


class _TextureGaussians(Function):
    """Host side of texture_forward_tensor / texture_backward_tensor (reference texture.py:156-408) over the staged C-ABI
    entry points: pack -> pad -> [intersection count read back] -> fused tile binning -> rasterise forward, and
    rasterise backward -> per-Gaussian epilogue -> un-pad.  The one host synchronisation the reference has per call
    (``num_intersects = cum_tiles_hit[-1].item()``, utils.py:58) is kept - the output sizes depend on it - but the copy
    is asynchronous and everything that does not depend on the count (output allocation, record packing, texture
    padding) is enqueued before the host waits for it, so the GPU keeps working through the round trip."""

    @staticmethod
    def forward(ctx, texture_info, texture_dims, centers, extents, depths, num_tiles_hit, colors, opacity, means,
                scales, glob_scale, quats, uv0, umap, vmap, texture, viewmat, c2w, fx, fy, cx, cy, img_height,
                img_width, block_width, settings, background, max_intersects=None, texture_grad=None):
        lib = _lib.load()
        H, W, bw = int(img_height), int(img_width), int(block_width)
        tile_bounds = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
        dev = centers.device
        C = int(texture_info[2])
        n, X = means.shape[0], texture.shape[0]
        _C._check_raster_inputs(texture_dims, None, None, colors, opacity, means, scales, quats, uv0, umap, vmap, texture,
                                viewmat, c2w, background)
        # (X,4) texels with texture_info[2] == 3: the three channels stored at a 16-byte pitch - read, and differentiated,
        # in place (no padded copy, no un-padding of the gradient); an opt-in layout beyond the reference's (X,C)
        rgba = C == 3 and texture.dim() == 2 and texture.shape[1] == 4
        if texture.dim() != 2 or (texture.shape[1] != C and not rgba):
            raise RuntimeError(f"texture must have dimensions (X, {C})")
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev)
        s = stream.cuda_stream
        fx, fy, cx, cy = float(fx), float(fy), float(cx), float(cy)
        with torch.cuda.device(dev):
            # 1. inclusive scan of the tile counts; its last element starts travelling to the host
            count_host, count_ready = None, None
            if n > 0 and max_intersects is None:
                cum = cumsum_i32(num_tiles_hit.reshape(-1))
                count_host = torch.empty((1,), dtype=torch.int32, pin_memory=True)
                count_host.copy_(cum[-1:], non_blocking=True)
                count_ready = torch.cuda.Event()
                count_ready.record(stream)
            # 2. everything that does not depend on the count
            out_img, out_depth, out_reg = torch.empty((H, W, 3), **f32), torch.empty((H, W), **f32), torch.empty((H, W), **f32)
            out_texture, out_normal = torch.empty((H, W, C), **f32), torch.empty((H, W, 3), **f32)
            final_Ts, final_idx, depth_idx = torch.empty((H, W), **f32), torch.empty((H, W), **i32), torch.empty((H, W), **i32)
            out_reg_s = torch.empty((H, W, 3), **f32)
            recs, mean2d = torch.empty((max(n, 1), 32), **f32), torch.empty((max(n, 1), 2), **f32)
            tex4 = texture if rgba else (torch.empty((max(X, 1), 4), **f32) if C == 3 else None)
            if n > 0:
                _lib.check(lib.gstex_pack_records(n, _p(texture_dims), _p(colors), _p(opacity), _p(means), _p(scales),
                                                  float(glob_scale), _p(quats), _p(uv0), _p(umap), _p(vmap), _p(viewmat),
                                                  _p(c2w), fx, fy, cx, cy, _p(recs), _p(mean2d), 0, s), "pack_records")
            if C == 3 and X > 0 and not rgba:
                _lib.check(lib.gstex_pad_texture(X, _p(texture), _p(tex4), s), "pad_texture")
            # 3. the count (the reference's one sync per call)
            num_intersects = 0
            if count_ready is not None:
                count_ready.synchronize()
                num_intersects = int(count_host[0])
        if max_intersects is not None:
            num_intersects = max(int(max_intersects), 1)  # a capacity: the true count stays on the device
        ctx.num_intersects = num_intersects
        ctx.texture_grad = texture_grad
        if num_intersects < 1:
            # upstream leaves several outputs undefined in this branch (texture.py:197-205, :254-289);
            # we return the background-only image and zeros
            out_img = torch.ones(H, W, colors.shape[-1], **f32) * background
            zeros = torch.zeros(H, W, **f32)
            ctx.save_for_backward(colors, opacity, means, scales, quats, uv0, umap, vmap, texture)
            return (out_img, zeros, zeros.clone(), zeros.clone(), torch.zeros(H, W, C, **f32), torch.zeros(H, W, 3, **f32))
        # 4. same gaussian_ids_sorted / tile_bins as bin_and_sort_gaussians (utils.py:106-162 upstream), from the fused
        #    bucket-by-tile + per-tile sort (csrc/binning_tiles.cu) instead of the global 64-bit key sort
        gaussian_ids_sorted, tile_bins, count_dev, _ = bin_tiles(centers, extents, depths, tile_bounds, bw, num_intersects)
        if max_intersects is not None:
            _LAST_COUNT[dev.index] = count_dev
        # blend masks: forward -> backward (csrc/raster.cuh); not kept (nor zero-filled) for an inference-only call
        need_grad = any(ctx.needs_input_grad)
        masks = torch.empty((num_intersects, 8), **i32) if need_grad else None
        tex = tex4 if C == 3 else texture
        with torch.cuda.device(dev):
            rc = lib.gstex_raster_forward(H, W, bw, C, int(settings), _p(gaussian_ids_sorted), _p(tile_bins), _p(recs),
                                          _p(mean2d), _p(tex), _p(viewmat), _p(c2w), fx, fy, cx, cy, _p(background),
                                          _p(out_img), _p(out_depth), _p(out_reg), _p(out_texture), _p(out_normal),
                                          _p(final_Ts), _p(final_idx), _p(depth_idx), _p(out_reg_s), _p(masks),
                                          num_intersects, _p(count_dev) if max_intersects is not None else 0, s)
        _lib.check(rc, "raster_forward")
        ctx.img_width, ctx.img_height, ctx.block_width = W, H, bw
        ctx.texture_info, ctx.settings, ctx.glob_scale = texture_info, int(settings), float(glob_scale)
        ctx.intr = (fx, fy, cx, cy)
        if need_grad:
            ctx.save_for_backward(gaussian_ids_sorted, tile_bins, means, scales, quats, umap, vmap, texture, viewmat, c2w,
                                  background, final_Ts, final_idx, depth_idx, out_reg_s, recs, mean2d, tex4, masks, uv0)
        out_alpha = 1 - final_Ts
        return out_img, out_depth, out_reg, out_alpha, out_texture, out_normal

    @staticmethod
    def backward(ctx, v_out_img, v_out_depth, v_out_reg, v_out_alpha, v_out_texture, v_out_normal):
        none18 = [None] * 29
        fused_tex = ctx.texture_grad  # texel gradients are added into this buffer instead of being returned
        if ctx.num_intersects < 1:
            colors, opacity, means, scales, quats, uv0, umap, vmap, texture = ctx.saved_tensors
            grads = [torch.zeros_like(t) for t in (colors, opacity, means, scales, quats, uv0, umap, vmap)]
            grads.append(None if fused_tex is not None else torch.zeros_like(texture))
        else:
            (gaussian_ids_sorted, tile_bins, means, scales, quats, umap, vmap, texture, viewmat, c2w, background,
             final_Ts, final_idx, depth_idx, out_reg_s, recs, mean2d, tex4, masks, uv0) = ctx.saved_tensors
            lib = _lib.load()
            H, W, bw = ctx.img_height, ctx.img_width, ctx.block_width
            dev = means.device
            f32 = dict(dtype=torch.float32, device=dev)

            def dense(v, shape, needed=False):
                # an output the loss does not use has no upstream gradient: the kernel takes NULL for zeros
                if v is None:
                    return torch.zeros(shape, **f32) if needed else None
                return v.contiguous()

            C = int(ctx.texture_info[2])
            n, X = means.shape[0], texture.shape[0]
            fx, fy, cx, cy = ctx.intr
            v_img, v_dep, v_reg = dense(v_out_img, (H, W, 3)), dense(v_out_depth, (H, W)), dense(v_out_reg, (H, W))
            v_alp, v_nrm = dense(v_out_alpha, (H, W)), dense(v_out_normal, (H, W, 3))
            v_tex = dense(v_out_texture, (H, W, C), needed=C != 3)  # the generic-channel path reads it unconditionally
            for name, t in (("v_output", v_img), ("v_output_depth", v_dep), ("v_output_reg", v_reg),
                            ("v_output_alpha", v_alp), ("v_output_texture", v_tex), ("v_output_normal", v_nrm)):
                if t is not None:
                    _C._chk(name, t, torch.float32)
            acc = torch.zeros((n, 32), **f32)                       # moment lines (csrc/common.cuh: AccSlot)
            rgba = C == 3 and texture.shape[1] == 4
            if fused_tex is not None and (rgba or C != 3):
                vtex4 = v_texture = fused_tex  # the rasteriser's reductions land in the caller's buffer
            else:
                vtex4 = torch.zeros((X, 4), **f32) if C == 3 else None  # texel gradients at a 16-byte pitch
                v_texture = vtex4 if rgba else (torch.zeros((X, C), **f32) if C != 3 else torch.empty((X, C), **f32))
                if fused_tex is not None:
                    v_texture = fused_tex  # (X,3): the un-padding pass adds into it
            v_colors, v_opacity = torch.empty((n, 3), **f32), torch.empty((n, 1), **f32)
            v_means, v_scales, v_quats = torch.empty((n, 3), **f32), torch.empty((n, 3), **f32), torch.empty((n, 4), **f32)
            # shaped like the inputs ((n, num_probs, k) upstream, zero-filled: texture.cu:1004-1006); the kernels address
            # them flat, Gaussian g at element g, exactly as the reference kernels do
            if uv0.numel() == 2 * n:
                v_uv0, v_umap, v_vmap = torch.empty_like(uv0), torch.empty_like(umap), torch.empty_like(vmap)
            else:
                v_uv0, v_umap, v_vmap = torch.zeros_like(uv0), torch.zeros_like(umap), torch.zeros_like(vmap)
            s = torch.cuda.current_stream(dev).cuda_stream
            tex = tex4 if C == 3 else texture
            vtex = vtex4 if C == 3 else v_texture
            with torch.cuda.device(dev):
                rc = lib.gstex_raster_backward(H, W, bw, C, ctx.settings, _p(gaussian_ids_sorted), _p(tile_bins), _p(recs),
                                               _p(mean2d), _p(tex), _p(viewmat), _p(c2w), fx, fy, cx, cy, _p(background),
                                               _p(final_Ts), _p(final_idx), _p(depth_idx), _p(out_reg_s), _p(v_img),
                                               _p(v_dep), _p(v_reg), _p(v_alp), _p(v_tex), _p(v_nrm), _p(masks), _p(acc),
                                               _p(vtex), s)
                _lib.check(rc, "raster_backward")
                rc = lib.gstex_raster_epilogue(n, _p(means), _p(scales), ctx.glob_scale, _p(quats), _p(umap), _p(vmap),
                                               _p(viewmat), _p(c2w), fx, fy, cx, cy, _p(acc), _p(recs), _p(v_colors), _p(v_opacity),
                                               _p(v_means), _p(v_scales), _p(v_quats), _p(v_uv0), _p(v_umap), _p(v_vmap),
                                               0, s)
                _lib.check(rc, "raster_epilogue")
                if C == 3 and not rgba:
                    _lib.check(lib.gstex_unpad_texture_grad(X, _p(vtex4), _p(v_texture), 1 if fused_tex is not None else 0, s),
                               "unpad_texture_grad")
            grads = (v_colors, v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap,
                     None if fused_tex is not None else v_texture)
        v_colors, v_opacity, v_means, v_scales, v_quats, v_uv0, v_umap, v_vmap, v_texture = grads
        none18[6], none18[7], none18[8], none18[9] = v_colors, v_opacity, v_means, v_scales
        none18[11], none18[12], none18[13], none18[14], none18[15] = v_quats, v_uv0, v_umap, v_vmap, v_texture
        return tuple(none18)
