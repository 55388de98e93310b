"""Fused image loss of the reference trainer (example.py:189-209) as ONE kernel with an autograd wrapper:

    loss = mse(out_texture, gt) + out_reg.mean() + (nx^2 + ny^2 + (1 - nz)^2).mean()

The reference evaluates it with ~20 torch ops and autograd differentiates it with ~40 more, each a launch over a
full-resolution image; csrc/loss.cu produces the value and the three gradients in one pass.
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib


def image_loss(out_texture: Tensor, out_reg: Tensor, out_normal: Tensor, gt: Tensor) -> Tensor:
    """Scalar loss of example.py:189-209 for 3-channel ``out_texture`` / ``gt`` (H,W,3), ``out_reg`` (H,W) and
    ``out_normal`` (H,W,3).  Differentiable w.r.t. the three rasteriser outputs."""
    return _ImageLoss.apply(out_texture, out_reg, out_normal, gt)


class _ImageLoss(Function):
    @staticmethod
    def forward(ctx, out_texture: Tensor, out_reg: Tensor, out_normal: Tensor, gt: Tensor):
        for name, t in (("out_texture", out_texture), ("out_reg", out_reg), ("out_normal", out_normal), ("gt", gt)):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError(f"image_loss: {name} must be a float32 CUDA tensor (there is no CPU path)")
        H, W = out_reg.shape[-2], out_reg.shape[-1]
        if tuple(out_texture.shape) != (H, W, 3) or tuple(out_normal.shape) != (H, W, 3) or tuple(gt.shape) != (H, W, 3):
            raise RuntimeError("image_loss expects out_texture / out_normal / gt of shape (H, W, 3) and out_reg (H, W)")
        out_texture, out_reg, out_normal, gt = (t.contiguous() for t in (out_texture, out_reg, out_normal, gt))
        dev = out_reg.device
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        v_tex, v_nrm = torch.empty_like(out_texture), torch.empty_like(out_normal)
        v_reg = torch.empty_like(out_reg)
        # the three gradients this loss does not use (image, depth, alpha) are not wanted: NULL
        with torch.cuda.device(dev):
            rc = _lib.load().gstex_image_loss(H, W, out_texture.data_ptr(), out_reg.data_ptr(), out_normal.data_ptr(),
                                              gt.data_ptr(), loss.data_ptr(), 0, 0, v_reg.data_ptr(), 0,
                                              v_tex.data_ptr(), v_nrm.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "image_loss")
        ctx.save_for_backward(v_tex, v_reg, v_nrm)
        return loss

    @staticmethod
    def backward(ctx, v_loss: Tensor):
        v_tex, v_reg, v_nrm = ctx.saved_tensors
        return v_tex * v_loss, v_reg * v_loss, v_nrm * v_loss, None
