"""Times the UNMODIFIED reference CUDA extension (oracle/_ref) on the C4 workload on this GPU, next to our path.
Not part of bench.py's contract: it produces the "reference CUDA beside it" numbers of SURVEY 8d for DESIGN.md /
profiles/.  Usage (GPU box):  python tools/time_reference_cuda.py [points] [width] [height]
"""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gstex_cuda_b200.scenes import synthetic_scene  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
DEV = "cuda:0"

spec = importlib.util.spec_from_file_location("gstex_ref_C", os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

SCALE_MULT = float(os.environ.get("GSTEX_SCALE_MULT", "1.0"))  # bench.py --scale-mult
s = synthetic_scene(N, W, H, seed=1234, device=DEV, scale_lo=0.004 * SCALE_MULT, scale_hi=0.04 * SCALE_MULT)
fx, fy, cx, cy = s["intrins"]
bw = 16
tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
vm, c2w = s["viewmat"], s["c2w"]
gt = s["target"]
P = H * W


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step(times):
    """reference pipeline: example.py:146-152 + texture.py:195-289 + example.py:189-209, torch glue as upstream"""
    t = [ev()]
    dirs = (s["means"] - c2w[:3, 3]).contiguous()
    colors = torch.clamp(ref.compute_sh_forward(N, 3, 3, dirs, s["sh_coeffs"]) + 0.5, 0, 1).contiguous()
    t.append(ev())
    depths = (s["means"] @ vm[:3, :3].T + vm[:3, 3])[:, 2].contiguous()
    centers, extents = ref.get_aabb_2d(s["means"], s["scales"], 1.0, s["quats"], vm, fx, fy, cx, cy)
    tl = torch.floor((centers - extents) / bw).to(torch.int32)
    br = torch.floor((centers + extents) / bw + 1).to(torch.int32)
    tmin = torch.stack([tl[:, 0].clamp(0, tb[0]), tl[:, 1].clamp(0, tb[1])], -1)
    tmax = torch.stack([br[:, 0].clamp(0, tb[0]), br[:, 1].clamp(0, tb[1])], -1)
    nth = ((tmax - tmin)[:, 0] * (tmax - tmin)[:, 1]).to(torch.int32)
    t.append(ev())
    cum = torch.cumsum(nth, 0, dtype=torch.int32)
    m = int(cum[-1].item())
    isect, gids = ref.map_gaussian_to_intersects(N, m, centers, extents, depths, cum, tb, bw, False)
    isect_s, perm = torch.sort(isect)
    gids_s = torch.gather(gids, 0, perm)
    bins = ref.get_tile_bin_edges(m, isect_s, tb)
    t.append(ev())
    outs = ref.texture_forward(tb, (bw, bw, 1), (W, H, 1), (N, 1, 3), s["texture_dims"], gids_s, bins, colors,
                               s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"],
                               s["texture"], vm, c2w, fx, fy, cx, cy, 1 << 8, s["background"])
    t.append(ev())
    out_tex, out_reg, out_n = outs[3], outs[2], outs[4]
    v_tex = (2.0 / (3 * P)) * (out_tex - gt)
    v_n = torch.stack([2 * out_n[..., 0], 2 * out_n[..., 1], -2 * (1 - out_n[..., 2])], -1) / P
    z = torch.zeros(H, W, device=DEV)
    v_reg = torch.full((H, W), 1.0 / P, device=DEV)
    t.append(ev())
    g = ref.texture_backward(H, W, bw, (N, 1, 3), s["texture_dims"], gids_s, bins, colors, s["opacities"], s["means"],
                             s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"], vm, c2w, fx, fy,
                             cx, cy, 1 << 8, s["background"], outs[5], outs[6], outs[7], outs[8],
                             torch.zeros(H, W, 3, device=DEV), z, v_reg, z, v_tex.contiguous(), v_n.contiguous())
    t.append(ev())
    ref.compute_sh_backward(N, 3, 3, dirs, g[0].contiguous())
    t.append(ev())
    torch.cuda.synchronize()
    names = ["sh_fwd", "project+aabb+count", "bin+sort", "raster_fwd", "loss_grad", "raster_bwd", "sh_bwd"]
    for n_, a, b in zip(names, t[:-1], t[1:]):
        times.setdefault(n_, []).append(a.elapsed_time(b))
    times.setdefault("total", []).append(t[0].elapsed_time(t[-1]))
    return m


for _ in range(3):
    step({})
times = {}
for _ in range(10):
    m = step(times)
avg = {k: sum(v) / len(v) for k, v in times.items()}
out = {"what": "UNMODIFIED reference CUDA extension (gstex_cuda @ abdc217, -O3, sm_100) on this B200", "points": N,
       "width": W, "height": H, "intersections": m, "ms": avg, "mpixel_per_s": H * W / (avg["total"] * 1e-3) / 1e6}
print(json.dumps(out))
