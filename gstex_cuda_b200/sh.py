"""Spherical-harmonics colour evaluation (mirror of ``gstex_cuda/sh.py``, victor-rong/GStex_cuda)."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib
from . import cuda as _C


def num_sh_bases(degree: int) -> int:
    """sh.py:10-19"""
    return {0: 1, 1: 4, 2: 9, 3: 16}.get(degree, 25)


def deg_from_sh(num_bases: int) -> int:
    """sh.py:22-33"""
    table = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}
    assert num_bases in table, "Invalid number of SH bases"
    return table[num_bases]


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """sh.py:36-57: colours = sum_b Y_b(dir / |dir|) * coeff[b]; differentiable w.r.t. coeffs only."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    return _SphericalHarmonics.apply(degrees_to_use, viewdirs.contiguous(), coeffs.contiguous())


class _SphericalHarmonics(Function):
    @staticmethod
    def forward(ctx, degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor):
        num_points = coeffs.shape[0]
        ctx.degrees_to_use = degrees_to_use
        ctx.degree = deg_from_sh(coeffs.shape[-2])
        ctx.save_for_backward(viewdirs)
        return _C.compute_sh_forward(num_points, ctx.degree, degrees_to_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        viewdirs = ctx.saved_tensors[0]
        return (None, None, _C.compute_sh_backward(v_colors.shape[0], ctx.degree, ctx.degrees_to_use, viewdirs,
                                                   v_colors.contiguous()))


def spherical_harmonics_colors(degrees_to_use: int, means: Tensor, c2w: Tensor, coeffs: Tensor,
                               coeffs_grad: Tensor = None) -> Tensor:
    """View-dependent colours as a trainer forms them around the SH op, in one kernel each way:
    ``clamp(spherical_harmonics(deg, means - c2w[:3, 3], coeffs) + 0.5, 0, 1)``.  Differentiable w.r.t. ``coeffs``
    only (like the reference op, sh.py:60-96, whose ``viewdirs`` carry no gradient); the clamp gates the gradient.

    ``coeffs_grad`` (opt-in): a float32 tensor shaped like ``coeffs`` that the backward pass ADDS the coefficient
    gradients into instead of returning them to autograd (see ``texture_gaussians(..., texture_grad=)``)."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    if coeffs_grad is not None and not (coeffs_grad.is_cuda and coeffs_grad.dtype == torch.float32
                                        and coeffs_grad.is_contiguous() and coeffs_grad.shape == coeffs.shape
                                        and coeffs_grad.device == coeffs.device):
        raise ValueError("coeffs_grad must be a contiguous float32 CUDA tensor shaped like coeffs")
    return _SphericalHarmonicsColors.apply(degrees_to_use, means.detach().contiguous(), c2w.contiguous(),
                                           coeffs.contiguous(), coeffs_grad)


class _SphericalHarmonicsColors(Function):
    @staticmethod
    def forward(ctx, degrees_to_use: int, means: Tensor, c2w: Tensor, coeffs: Tensor, coeffs_grad=None):
        ctx.coeffs_grad = coeffs_grad
        for name, t in (("means", means), ("c2w", c2w), ("coeffs", coeffs)):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError(f"spherical_harmonics_colors: {name} must be a float32 CUDA tensor")
        n, dev = coeffs.shape[0], coeffs.device
        ctx.degrees_to_use, ctx.degree = degrees_to_use, deg_from_sh(coeffs.shape[-2])
        colors = torch.empty((n, 3), dtype=torch.float32, device=dev)
        mask = torch.empty((n,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().gstex_sh_colors_forward(n, ctx.degree, degrees_to_use, means.data_ptr(), c2w.data_ptr(),
                                                     coeffs.data_ptr(), colors.data_ptr(), mask.data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "sh_colors_forward")
        ctx.save_for_backward(means, c2w, mask)
        ctx.coeff_shape = tuple(coeffs.shape)
        return colors

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        means, c2w, mask = ctx.saved_tensors
        n, dev = means.shape[0], means.device
        v_colors = v_colors.contiguous()
        fused = ctx.coeffs_grad
        # rows past degrees_to_use come back zero (or, when adding into the caller's buffer, stay untouched)
        # (the kernel writes every row, zeros past degrees_to_use: no zero-fill needed)
        v_coeffs = fused if fused is not None else torch.empty(ctx.coeff_shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().gstex_sh_colors_backward(n, ctx.degree, ctx.degrees_to_use, means.data_ptr(), c2w.data_ptr(),
                                                      v_colors.data_ptr(), mask.data_ptr(), v_coeffs.data_ptr(),
                                                      1 if fused is not None else 0,
                                                      torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "sh_colors_backward")
        return None, None, None, (None if fused is not None else v_coeffs), None
