"""Deterministic synthetic scenes for tests and benchmarks.

All random numbers come from a CPU ``torch.Generator`` so that the same seed gives the same scene on
every device; tensors are moved to ``device`` at the end.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch


def look_at_camera(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """World->camera 4x4 (camera looks down +z, x right, y down) and its inverse."""
    eye = torch.tensor(eye, dtype=torch.float64)
    fwd = torch.tensor(target, dtype=torch.float64) - eye
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float64)
    right = torch.linalg.cross(fwd, upv)
    if right.norm() < 1e-8:
        right = torch.linalg.cross(fwd, torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], dim=0)  # rows: camera axes in world coordinates
    viewmat = torch.eye(4, dtype=torch.float64)
    viewmat[:3, :3] = R
    viewmat[:3, 3] = -R @ eye
    c2w = torch.linalg.inv(viewmat)
    # torch.linalg.inv returns a column-major tensor: make both row-major (the kernels read raw pointers)
    return viewmat.float().contiguous(), c2w.float().contiguous()


def uniform_quats(n: int, gen: torch.Generator) -> torch.Tensor:
    """Shoemake-uniform unit quaternions (example.py:77-88)."""
    u, v, w = torch.rand(n, 1, generator=gen), torch.rand(n, 1, generator=gen), torch.rand(n, 1, generator=gen)
    q = torch.cat([torch.sqrt(1 - u) * torch.sin(2 * math.pi * v), torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
                   torch.sqrt(u) * torch.sin(2 * math.pi * w), torch.sqrt(u) * torch.cos(2 * math.pi * w)], -1)
    return q / q.norm(dim=-1, keepdim=True)


def quat_axes(q: torch.Tensor):
    w, x, y, z = q.unbind(-1)
    a1 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], -1)
    a2 = torch.stack([2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)], -1)
    a3 = torch.stack([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)], -1)
    return a1, a2, a3


def synthetic_scene(num_points: int, width: int, height: int, *, seed: int = 1234, th: int = 4, tw: int = 4,
                    channels: int = 3, sh_degree: int = 3, scale_lo: float = 0.004, scale_hi: float = 0.04,
                    device="cpu", background=(0.0, 0.0, 0.0)) -> Dict[str, object]:
    """Config C4: Gaussians back-projected from uniformly random pixels at view depth U(6,10)
    in front of a camera at z = -8 looking down +z with a 90 degree horizontal field of view."""
    gen = torch.Generator().manual_seed(seed)
    N, W, H = num_points, width, height
    f = 0.5 * W / math.tan(0.25 * math.pi)
    cx, cy = W / 2.0, H / 2.0
    viewmat = torch.eye(4)
    viewmat[2, 3] = 8.0
    c2w = torch.linalg.inv(viewmat)
    u = torch.rand(N, generator=gen) * W
    v = torch.rand(N, generator=gen) * H
    d = 6.0 + 4.0 * torch.rand(N, generator=gen)
    means = torch.stack([(u - cx) * d / f, (v - cy) * d / f, d - 8.0], -1)
    ls = math.log(scale_lo) + (math.log(scale_hi) - math.log(scale_lo)) * torch.rand(N, 2, generator=gen)
    s12 = torch.exp(ls)
    scales = torch.cat([s12, 1e-5 * s12.mean(dim=-1, keepdim=True)], -1)
    quats = uniform_quats(N, gen)
    opacities = torch.sigmoid(1.5 * torch.randn(N, 1, generator=gen))
    K = (sh_degree + 1) ** 2
    sh = 0.1 * torch.randn(N, K, 3, generator=gen)
    sh[:, 0, :] = torch.rand(N, 3, generator=gen) / 0.28209479177387814
    texture = torch.rand(N * th * tw, channels, generator=gen)
    dims = torch.zeros(N, 3, dtype=torch.int32)
    dims[:, 0], dims[:, 1] = th, tw
    dims[:, 2] = torch.arange(N, dtype=torch.int32) * (th * tw)
    a1, a2, _ = quat_axes(quats)
    uv0 = torch.full((N, 1, 2), 0.5)
    umap = (a1 / (6.0 * s12[:, :1]))[:, None, :]
    vmap = (a2 / (6.0 * s12[:, 1:2]))[:, None, :]
    target = torch.rand(H, W, 3, generator=gen)
    out = dict(
        means=means, scales=scales, quats=quats, opacities=opacities, sh_coeffs=sh, texture=texture,
        texture_dims=dims, uv0=uv0, umap=umap, vmap=vmap, viewmat=viewmat, c2w=c2w, target=target,
        background=torch.tensor(background, dtype=torch.float32),
    )
    out = {k: (t.contiguous().to(device) if torch.is_tensor(t) else t) for k, t in out.items()}
    out.update(intrins=(f, f, cx, cy), H=H, W=W, num_points=N, texture_info=(N, 1, channels), sh_degree=sh_degree,
               glob_scale=1.0, settings=1 << 8, block_width=16)
    return out


def circle_cameras(num_views: int, radius: float = 8.0, height: float = 0.0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Config C5: cameras on a circle around the origin, all looking at it (deterministic)."""
    cams = []
    for i in range(num_views):
        a = 2.0 * math.pi * i / num_views
        eye = (radius * math.sin(a), height, -radius * math.cos(a))
        cams.append(look_at_camera(eye))
    return cams


def arc_cameras(num_views: int, radius: float = 8.0, half_angle_deg: float = 30.0) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Config C5 as benchmarked: cameras on an arc of +-half_angle around the C4 front camera (same distance, all
    looking at the origin), so that every view sees the scene face-on and costs about what the C4 view costs -
    which is what makes the 1 -> 8 GPU throughput numbers comparable.  View num_views // 2 is (nearly) the front view."""
    cams = []
    for i in range(num_views):
        a = math.radians(half_angle_deg) * (2.0 * i / max(1, num_views - 1) - 1.0) if num_views > 1 else 0.0
        cams.append(look_at_camera((radius * math.sin(a), 0.0, -radius * math.cos(a))))
    return cams


def random_small_scene(num_points: int, width: int, height: int, *, seed: int = 0, th: int = 5, tw: int = 3,
                       channels: int = 3, device="cpu", spread: float = 16.0, scale_pow: float = 0.5,
                       background=(0.1, 0.4, 0.7), jagged: bool = True) -> Dict[str, object]:
    """Small test scene in the style of example.py:69-119 (means in a slab in front of a camera at z=-8),
    with jagged texture sizes, in-plane uv maps of random scale / rotation and random opacities."""
    gen = torch.Generator().manual_seed(seed)
    N, W, H = num_points, width, height
    f = 0.5 * W / math.tan(0.25 * math.pi)
    means = torch.rand(N, 3, generator=gen) - 0.5
    means[:, :2] *= spread
    means[:, 0] *= W / max(W, H)
    means[:, 1] *= H / max(W, H)
    s12 = torch.exp(scale_pow * math.log(1.0 / N) * torch.rand(N, 2, generator=gen))
    scales = torch.cat([s12, 1e-5 * s12.mean(dim=-1, keepdim=True)], -1)
    quats = uniform_quats(N, gen)
    opacities = torch.sigmoid(2.0 * torch.randn(N, 1, generator=gen))
    colors = torch.rand(N, 3, generator=gen)
    if jagged:
        hs = torch.randint(1, th + 1, (N,), generator=gen, dtype=torch.int32)
        ws = torch.randint(1, tw + 1, (N,), generator=gen, dtype=torch.int32)
    else:
        hs = torch.full((N,), th, dtype=torch.int32)
        ws = torch.full((N,), tw, dtype=torch.int32)
    dims = torch.stack([hs, ws, torch.cumsum(hs * ws, 0).to(torch.int32) - hs * ws], -1).contiguous()
    texture = torch.rand(int((hs * ws).sum()), channels, generator=gen)
    a1, a2, _ = quat_axes(quats)
    theta = 2 * math.pi * torch.rand(N, 1, generator=gen)
    uvs = torch.exp(torch.randn(N, 1, generator=gen) * 0.3) / (5.0 * s12.mean(dim=-1, keepdim=True))
    umap = (uvs * (a1 * torch.cos(theta) + a2 * torch.sin(theta)))[:, None, :]
    vmap = (uvs * (-a1 * torch.sin(theta) + a2 * torch.cos(theta)))[:, None, :]
    uv0 = 0.5 + 0.1 * torch.randn(N, 1, 2, generator=gen)
    viewmat = torch.eye(4)
    viewmat[2, 3] = 8.0
    out = dict(means=means, scales=scales, quats=quats, opacities=opacities, colors=colors, texture=texture,
               texture_dims=dims, uv0=uv0, umap=umap, vmap=vmap, viewmat=viewmat, c2w=torch.linalg.inv(viewmat),
               background=torch.tensor(background, dtype=torch.float32), target=torch.rand(H, W, 3, generator=gen))
    out = {k: (t.contiguous().to(device) if torch.is_tensor(t) else t) for k, t in out.items()}
    out.update(intrins=(f, f, W / 2.0, H / 2.0), H=H, W=W, num_points=N, texture_info=(N, 1, channels),
               glob_scale=1.0, settings=1 << 8, block_width=16)
    return out
