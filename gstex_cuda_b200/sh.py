"""Spherical-harmonics colour evaluation (mirror of ``gstex_cuda/sh.py``, victor-rong/GStex_cuda)."""
from __future__ import annotations

from torch import Tensor
from torch.autograd import Function

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
