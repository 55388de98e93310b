"""Projection, screen-space AABB and tile counts.

Mirror of ``gstex_cuda/get_aabb_2d.py`` (victor-rong/GStex_cuda).  ``project_points`` and
``get_num_tiles_hit_2d`` are torch-op chains upstream (get_aabb_2d.py:22-32, :70-92); here each is one
CUDA kernel of libgstex_b200 (csrc/project.cu).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib
from . import cuda as _C
from . import _torch_impl as _T


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def merge_aabbs(centers1, extents1, centers2, extents2):
    """get_aabb_2d.py:15-20"""
    lower = torch.minimum(centers1 - extents1, centers2 - extents2)
    upper = torch.maximum(centers1 + extents1, centers2 + extents2)
    return 0.5 * (lower + upper), 0.5 * (upper - lower)


class _ProjectPoints(Function):
    @staticmethod
    def forward(ctx, points, viewmat, fx, fy, cx, cy):
        pts = points.contiguous().float()
        vm = viewmat.contiguous().float()
        if vm.numel() < 12:
            raise RuntimeError("viewmat must hold at least the first three rows of a 4x4 matrix")
        n = pts.shape[0]
        pix = torch.empty((n, 2), dtype=torch.float32, device=pts.device)
        depths = torch.empty((n,), dtype=torch.float32, device=pts.device)
        with torch.cuda.device(pts.device):
            rc = _lib.load().gstex_project_points(n, pts.data_ptr(), vm.data_ptr(), float(fx), float(fy), float(cx),
                                                  float(cy), pix.data_ptr(), depths.data_ptr(), _stream(pts.device))
        _lib.check(rc, "project_points")
        ctx.save_for_backward(pts, vm)
        ctx.intr = (float(fx), float(fy))
        return pix, depths

    @staticmethod
    def backward(ctx, v_pix, v_depths):
        pts, vm = ctx.saved_tensors
        fx, fy = ctx.intr
        R, t = vm.reshape(-1)[:12].reshape(3, 4)[:, :3], vm.reshape(-1)[:12].reshape(3, 4)[:, 3]
        pv = pts @ R.T + t
        rw = 1.0 / (pv[:, 2] + 1e-6)
        gx, gy = fx * v_pix[:, 0], fy * v_pix[:, 1]
        v_pv = torch.stack([gx * rw, gy * rw, -(gx * pv[:, 0] + gy * pv[:, 1]) * rw * rw + v_depths], dim=-1)
        return v_pv @ R, None, None, None, None, None


def project_points(points, viewmat, intrins, clip=False):
    """get_aabb_2d.py:22-32: (pix, depths) with depths = view-space z (not clipped unless clip=True)."""
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor")
    if clip:
        # the reference clamps z in place under no_grad before projecting (get_aabb_2d.py:24-28)
        vm = viewmat.reshape(-1)[:12].reshape(3, 4)
        view_points = points @ vm[:, :3].T + vm[:, 3]
        with torch.no_grad():
            view_points[:, -1] = torch.clamp(view_points[:, -1], min=0.01, max=1000.0)
        return _T.project_pix(intrins[:2], view_points, intrins[2:]), view_points[:, -1]
    fx, fy, cx, cy = intrins
    return _ProjectPoints.apply(points, viewmat, fx, fy, cx, cy)


def get_aabb_2d_torch(means, scales, glob_scale, quats, viewmat, intrins):
    """get_aabb_2d.py:34-54: torch twin of the AABB kernel (corner z clamped to [0.01, 1000])."""
    ell = 3.0 * glob_scale
    Rs = _T.normalized_quat_to_rotmat(quats)
    a, b = ell * scales[:, None, 0] * Rs[:, :, 0], ell * scales[:, None, 1] * Rs[:, :, 1]
    corners = torch.stack([means + a + b, means + a - b, means - a + b, means - a - b], dim=1)
    shape = corners.shape
    proj = project_points(corners.reshape(-1, 3), viewmat, intrins, clip=True)[0].reshape(shape[0], shape[1], 2)
    hi, lo = proj.max(dim=1)[0], proj.min(dim=1)[0]
    return 0.5 * (hi + lo), 0.5 * (hi - lo)


def get_aabb_2d(means, scales, glob_scale, quats, viewmat, intrins):
    """get_aabb_2d.py:56-68"""
    fx, fy, cx, cy = intrins
    return _C.get_aabb_2d(means.contiguous(), scales.contiguous(), glob_scale, quats.contiguous(),
                          viewmat.contiguous(), fx, fy, cx, cy)


def get_num_tiles_hit_2d(centers, extents, img_height, img_width, block_width):
    """get_aabb_2d.py:70-92: number of tiles in the floor-based tile bbox of each AABB (int32)."""
    if not centers.is_cuda:
        raise RuntimeError("centers must be a CUDA tensor")
    c, e = centers.detach().contiguous().float(), extents.detach().contiguous().float()
    n = c.shape[0]
    out = torch.empty((n,), dtype=torch.int32, device=c.device)
    with torch.cuda.device(c.device):
        rc = _lib.load().gstex_num_tiles_hit_2d(n, c.data_ptr(), e.data_ptr(), int(img_height), int(img_width),
                                                int(block_width), out.data_ptr(), _stream(c.device))
    _lib.check(rc, "get_num_tiles_hit_2d")
    return out
