"""Bilinear queries into the jagged texture (mirror of ``gstex_cuda/texture_sample.py``)."""
from __future__ import annotations

from torch import Tensor
from torch.autograd import Function

from . import cuda as _C
from . import _torch_impl as _T


def texture_sample(texture_info, texture_dims, texture, uvs, use_torch_impl: bool = False,
                   texture_grad: bool = False) -> Tensor:
    """texture_sample.py:13-47: (Q, C) samples; uv clamped to [0,1]^2.

    Upstream's autograd backward returns None for every input (texture_sample.py:71-78) and that is the
    default here too.  ``texture_grad=True`` (not in the reference signature) lets the texture receive
    its gradient through the scatter kernel, the transpose of the fetch; uvs never get a gradient.
    """
    if use_torch_impl:
        return _T.sample_texture(texture_dims, texture, uvs)
    return _TextureSample.apply(texture_info, texture_dims.contiguous(), texture.contiguous(), uvs.contiguous(),
                                texture_grad)


class _TextureSample(Function):
    @staticmethod
    def forward(ctx, texture_info, texture_dims, texture, uvs, texture_grad=False) -> Tensor:
        out = _C.texture_sample_forward(texture_info, texture_dims, uvs, texture)
        ctx.texture_info = texture_info
        ctx.texture_grad = texture_grad
        ctx.save_for_backward(texture_dims, texture, uvs)
        return out

    @staticmethod
    def backward(ctx, v_out):
        texture_dims, texture, uvs = ctx.saved_tensors
        v_texture = None
        if ctx.texture_grad and ctx.needs_input_grad[2]:
            v_texture = _C.texture_sample_backward(ctx.texture_info, texture_dims, uvs, texture, v_out.contiguous())
        return None, None, v_texture, None, None
