#!/usr/bin/env python
"""Headline benchmark: forward+backward Mpixels/s of the textured-2DGS rasteriser hot path at 1080p with
1M textured Gaussians (BASELINE.json), on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm on the host cores (CPU oracle port)

A "step" is one training step of the hot path over one batch of synthetic views:
  SH colours -> project / AABB / tile count -> scan -> key emit -> radix sort -> tile ranges -> pack ->
  rasterise forward -> image loss (example.py:189-209) -> rasterise backward -> per-Gaussian epilogue ->
  SH backward [-> NCCL all-reduce of the gradient arena when N > 1].
N = 1 : BASELINE config 4 - one 1920x1080 view per step (the front camera of SURVEY 8d C4).
N > 1 : BASELINE config 5 - 64 views per step (an arc of +-30 degrees around the C4 camera, so that a view costs about
        what the C4 view costs), sharded round-robin over the ranks, replicated parameters, one all-reduce of the
        0.46 GB fp32 gradient arena per step  (strong scaling of the 64-view batch).
`value` = pixels rendered (forward+backward) by all ranks / device time (max over ranks), inputs resident in HBM.
`e2e`   = the same metric through the reference-shaped public API (spherical_harmonics_colors, project_points,
          get_aabb_2d, get_num_tiles_hit_2d, texture_gaussians, image_loss; torch autograd) with the step's inputs (camera matrices and target
          image) copied from pinned host memory and the loss read back, inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+bwd Mpixels/sec at 1080p, 1M textured Gaussians"
UNIT = "Mpixel/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--views", type=int, default=0, help="views per step (0: 1 at N=1, 64 at N>1)")
    ap.add_argument("--scale-mult", type=float, default=1.0,
                    help="multiplies the C4 scene's Gaussian scales (SURVEY 8d: a denser second data point at 2.0)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true",
                    help="skip timing the unmodified reference CUDA extension (oracle/_ref) beside our path at N=1")
    ap.add_argument("--no-scale-base", action="store_true",
                    help="skip the 64-view C5 batch at N=1 (the 1-GPU point of the scaling curve)")
    ap.add_argument("--texture-layout", default="rgba", choices=["rgba", "rgb"],
                    help="rgba: texels stored (X,4) at a 16-byte pitch, read / accumulated in place; rgb: (X,3), padded per step")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the frame in the CPU sample (0 = auto)")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_first(self, timeout_s: float = 15.0):
        """Block until nvidia-smi has delivered its first sample (it can take seconds, longer with several ranks
        querying at once): a short timed region must not end before the sampler has started."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout_s:
            time.sleep(0.05)

    def mark(self):
        """Index of the next sample: call at the start of the loaded region."""
        return len(self.rows)

    def stop(self, since: int = 0):
        """Summary of the samples taken since `since` (they cover warm-up + timed steps: all of it is load)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # let the last sample of the loaded region arrive
        self.proc.terminate()
        rows = self.rows[since:] or self.rows[-1:]
        sm, smax, reasons = [], None, set()
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 6:
                continue
            try:
                sm.append(float(c[0]))
                smax = float(c[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores, on a crop of the same workload
# --------------------------------------------------------------------------------------------------
def cpu_step(scene_np, rows: int):
    """One forward+backward of the top `rows` rows of the frame with the CPU oracle.  Returns seconds."""
    import numpy as np

    import oracle

    s = scene_np
    H, W, bw = rows, s["W"], 16
    fx, fy, cx, cy = s["intrins"]
    t0 = time.perf_counter()
    dirs = s["means"] - s["c2w"][:3, 3]
    raw = oracle.sh_forward(s["sh_degree"], s["sh_degree"], dirs, s["sh_coeffs"]) + 0.5
    colors = np.clip(raw, 0.0, 1.0).astype(np.float32)
    b = oracle.bin_view(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], s["intrins"], H, W, bw)
    args = (H, W, bw, s["texture_dims"], b["gaussian_ids_sorted"], b["tile_bins"], colors, s["opacities"], s["means"],
            s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"], s["viewmat"], s["c2w"], fx, fy,
            cx, cy, 1 << 8, s["background"])
    f = oracle.texture_forward(*args)
    P = H * W
    gt = s["target"][:H]
    v_tex = (2.0 * (f["out_texture"] - gt) / (3 * P)).astype(np.float32)
    n_ = f["out_normal"]
    v_n = (np.stack([2 * n_[..., 0], 2 * n_[..., 1], -2 * (1 - n_[..., 2])], -1) / P).astype(np.float32)
    z = np.zeros((H, W), np.float32)
    g = oracle.texture_backward(*args, f["final_Ts"], f["final_idx"], f["depth_idx"], f["out_reg_s"],
                                np.zeros((H, W, 3), np.float32), z, np.full((H, W), 1.0 / P, np.float32), z, v_tex, v_n)
    gate = ((raw > 0) & (raw < 1)).astype(np.float32)
    oracle.sh_backward(s["sh_degree"], s["sh_degree"], dirs, g["v_colors"] * gate)
    return time.perf_counter() - t0


def scene_to_numpy(scene):
    import torch

    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in scene.items()}


def run_reference(args, rank: int):
    """--impl reference: rank 0 times the CPU oracle (a port of the reference's algorithm; the reference's own
    CPU twin is Python and cannot travel to the GPU box) on a bounded crop, with all host threads."""
    if rank != 0:
        return
    import oracle
    from gstex_cuda_b200.scenes import synthetic_scene

    scene = scene_to_numpy(synthetic_scene(args.points, args.width, args.height, seed=1234))
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every host core regardless
    oracle.set_num_threads(os.cpu_count() or 1)
    cores = oracle.num_threads()
    # bounded sample: 128 rows of the frame cost about 1.1 s per step on 16 cores; shrink the crop for long runs so that
    # the whole --steps K run stays within about a minute (the metric is per pixel, the crop is named in `sample`)
    rows = args.cpu_rows or min(128, 128 * 40 // max(1, args.steps))
    rows = min(args.height, max(16, rows // 16 * 16))
    for _ in range(min(args.warmup, 1)):
        cpu_step(scene, rows)
    times = [cpu_step(scene, rows) for _ in range(max(1, args.steps))]
    t = sum(times) / len(times)
    value = rows * args.width / t / 1e6
    sample = f"top {rows} of {args.height} rows of the C4 frame ({args.width}x{rows} px), {args.points} Gaussians, fwd+bwd"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4 crop on host cores: " + sample, "gaussians": args.points, "sh_degree": 3,
                   "texels_per_gaussian": 16, "width": args.width, "height": rows},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# The unmodified reference CUDA extension beside our path (SURVEY 8d "Reference CUDA beside it"): a REPORTED comparison,
# timed after our timed region on the same GPU, same inputs.  oracle/_ref is checker infrastructure: nothing of it is on
# the measured path of `value` / `e2e`.
# --------------------------------------------------------------------------------------------------
def time_reference_cuda(scene, H, W, steps=5, warmup=2):
    import importlib.util

    import torch

    so = os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/gstex_ref_C.so not built (python oracle/build_ref.py in the build container)"}
    spec = importlib.util.spec_from_file_location("gstex_ref_C", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    s, N, bw = scene, scene["num_points"], 16
    dev = s["means"].device
    fx, fy, cx, cy = s["intrins"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    vm, c2w, gt, P = s["viewmat"], s["c2w"], s["target"], H * W

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(times):
        """the reference pipeline: example.py:146-152 + texture.py:195-289 + example.py:189-209, torch glue as upstream"""
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
                                   s["opacities"], s["means"], s["scales"], 1.0, s["quats"], s["uv0"], s["umap"],
                                   s["vmap"], s["texture"], vm, c2w, fx, fy, cx, cy, 1 << 8, s["background"])
        t.append(ev())
        out_tex, out_n = outs[3], outs[4]
        v_tex = (2.0 / (3 * P)) * (out_tex - gt)
        v_n = torch.stack([2 * out_n[..., 0], 2 * out_n[..., 1], -2 * (1 - out_n[..., 2])], -1) / P
        z = torch.zeros(H, W, device=dev)
        v_reg = torch.full((H, W), 1.0 / P, device=dev)
        t.append(ev())
        g = ref.texture_backward(H, W, bw, (N, 1, 3), s["texture_dims"], gids_s, bins, colors, s["opacities"], s["means"],
                                 s["scales"], 1.0, s["quats"], s["uv0"], s["umap"], s["vmap"], s["texture"], vm, c2w, fx,
                                 fy, cx, cy, 1 << 8, s["background"], outs[5], outs[6], outs[7], outs[8],
                                 torch.zeros(H, W, 3, device=dev), z, v_reg, z, v_tex.contiguous(), v_n.contiguous())
        t.append(ev())
        ref.compute_sh_backward(N, 3, 3, dirs, g[0].contiguous())
        t.append(ev())
        torch.cuda.synchronize()
        names = ["sh_fwd", "project+aabb+count", "bin+sort", "raster_fwd", "loss_grad", "raster_bwd", "sh_bwd"]
        for n_, a, b in zip(names, t[:-1], t[1:]):
            times.setdefault(n_, []).append(a.elapsed_time(b))
        times.setdefault("total", []).append(t[0].elapsed_time(t[-1]))
        return m

    for _ in range(warmup):
        step({})
    times = {}
    for _ in range(steps):
        m = step(times)
    avg = {k: sum(v) / len(v) for k, v in times.items()}
    return {"what": "UNMODIFIED reference CUDA extension (gstex_cuda @ abdc217, its JIT flags -O3, sm_100) through its own "
                    "Python-level pipeline, same GPU, same scene, after our timed region; reported comparison only",
            "steps": steps, "intersections": m, "ms": avg, "ms_per_step": avg["total"],
            "mpixel_per_s": H * W / (avg["total"] * 1e-3) / 1e6}


# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # rank 0 prints exactly one JSON line on stdout.  NCCL writes its banner and logs to stdout at any NCCL_DEBUG level
    # >= VERSION: when the caller asks for NCCL logs they are sent to stderr instead, so that the rank count can be read
    # from them without breaking the one-line contract.
    if os.environ.get("NCCL_DEBUG"):
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist

    from gstex_cuda_b200.pipeline import FusedTrainStep, DataParallelTrainStep
    from gstex_cuda_b200.scenes import synthetic_scene, arc_cameras

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi takes a while to deliver its first sample: start it long before the timed region
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    H, W, N = args.height, args.width, args.points

    scene = synthetic_scene(N, W, H, seed=1234, device=dev, scale_lo=0.004 * args.scale_mult,
                            scale_hi=0.04 * args.scale_mult)
    params = {k: scene[k] for k in ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")}
    rgba = args.texture_layout == "rgba"
    if rgba:  # the same texels at a 16-byte pitch: read and differentiated in place, no per-step padding passes
        params["texture"] = torch.cat([scene["texture"], torch.zeros_like(scene["texture"][:, :1])], 1).contiguous()
    views = args.views or (1 if world == 1 else 64)

    def make_batch(nviews):
        """cameras, host targets of this rank's shard, device targets, workload label"""
        if nviews == 1:
            cams = [(scene["viewmat"], scene["c2w"])]
            label = "C4: 1 view/step, front camera" + (f" (Gaussian scales x{args.scale_mult:g})" if args.scale_mult != 1.0 else "")
        else:
            cams = [(a.to(dev), b.to(dev)) for a, b in arc_cameras(nviews)]
            label = (f"C5: {nviews} views/step on a +-30 degree arc around the C4 camera, sharded over {world} GPU(s), "
                     f"one NCCL all-reduce of the gradient arena per step")
        mine = DataParallelTrainStep.shard(nviews, rank, world)
        g2 = torch.Generator().manual_seed(99)
        th = {}
        for v in range(nviews):
            t = torch.rand(H, W, 3, generator=g2)
            if v in mine:
                th[v] = t.pin_memory()
        tg = [th[v].to(dev) if v in th else None for v in range(nviews)]
        return cams, mine, th, tg, label

    cams, mine, targets_host, targets, workload = make_batch(views)
    fused = FusedTrainStep(params, scene["texture_dims"], H, W, intrins=scene["intrins"], sh_degree=scene["sh_degree"],
                           background=scene["background"], max_intersects=int(12 * N * max(1.0, args.scale_mult ** 2)),
                           texture_rgba=rgba)
    dp = DataParallelTrainStep(fused, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(cams_, targets_, nviews, steps, warmup, per_kernel):
        """W untimed steps, then exactly K steps between barrier + synchronize on both sides, CUDA events, max over ranks."""
        barrier()
        for _ in range(warmup):
            dp.step(cams_, targets_)
        barrier()
        m_max_ = fused.check_overflow()
        fused.time_kernels, fused.kernel_events = per_kernel, []
        dp.time_collective, dp.collective_events = world > 1, []
        launches0 = fused.launches
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss_ = dp.step(cams_, targets_)
        e1.record()
        barrier()
        el = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        lc = torch.tensor([fused.launches - launches0], device=dev, dtype=torch.float64)
        ar = [a.elapsed_time(b) for a, b in dp.collective_events]
        arm = torch.tensor([sum(ar) / len(ar) if ar else 0.0], device=dev, dtype=torch.float64)
        fused.time_kernels, dp.time_collective = False, False
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            dist.all_reduce(lc, op=dist.ReduceOp.SUM)
            dist.all_reduce(arm, op=dist.ReduceOp.MAX)
        ms = float(el.item()) / steps
        return dict(ms_per_step=ms, value=nviews * H * W / (ms * 1e-3) / 1e6, launches=int(lc.item()),
                    loss=float(loss_.item()), m_max=m_max_, allreduce_ms=float(arm.item()) if world > 1 else None)

    # ---- the headline run: W warm-up steps, exactly K timed steps ---------------------------------------------
    sampler.wait_first()
    load_start = sampler.mark()
    head = timed_run(cams, targets, views, args.steps, args.warmup, True)
    clocks = sampler.stop(load_start)
    ms_per_step, value = head["ms_per_step"], head["value"]

    # per-kernel device time inside the timed region (CUDA events on the launching stream)
    ktime = {}
    for name, a, b in fused.kernel_events:
        ktime.setdefault(name, []).append(a.elapsed_time(b))
    kavg = {k: sum(v) / len(v) for k, v in ktime.items()}
    m_view0 = int(fused.num_isect.item())  # intersections of the last view rendered by this rank

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peak, peak_src = hbm_peak()
    X, C, P = scene["texture"].shape[0], 3, H * W
    M = m_view0
    bytes_bwd = 104 * M + 12 * X + 72 * P + 100 * N + 12 * X      # SURVEY 8d, "raster bwd"
    bytes_fwd = 104 * M + 12 * X + 68 * P                          # SURVEY 8d, "raster fwd"
    bytes_step = 628 * N + 380 * M + 140 * P + 36 * X              # SURVEY 8d, whole step per view
    dom = max(("raster_backward", "raster_forward"), key=lambda k: kavg.get(k, 0.0))
    dom_bytes = bytes_bwd if dom == "raster_backward" else bytes_fwd
    dom_ms = kavg.get(dom, float("nan"))
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms == dom_ms and dom_ms > 0 else None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get(dom)
            traffic_src = "static: " + tj.get("_source", "profiles/dominant_kernel_traffic.json") + " - read from the committed capture, not measured in this run"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
                "share_of_step": (dom_ms * len(mine) / ms_per_step) if dom_ms == dom_ms else None,
                "step": {"algorithmic_bytes_per_view": bytes_step,
                         "achieved": bytes_step * len(mine) / (ms_per_step * 1e-3) / 1e9,
                         "frac": bytes_step * len(mine) / (ms_per_step * 1e-3) / 1e9 / peak},
                "kernel_ms": kavg}

    # ---- N = 1: the 64-view C5 batch as well, so that the 1 -> 8 GPU curve is ONE workload --------------------------
    scale_base = None
    if world == 1 and views == 1 and not args.no_scale_base and args.scale_mult == 1.0:
        c5 = make_batch(64)
        sb = timed_run(c5[0], c5[3], 64, 3, 1, False)
        scale_base = {"workload": c5[4], "views_per_step": 64, "steps": 3, "warmup": 1, "ms_per_step": sb["ms_per_step"],
                      "value": sb["value"], "unit": UNIT, "max_intersections_seen": sb["m_max"],
                      "note": "the N>1 lines of this bench run this 64-view batch (strong scaling); this is its 1-GPU point"}
        del c5

    # ---- end to end through the public (reference-shaped) API, host buffers -------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, scene, cams, mine, targets_host, dev, world, views)

    # ---- the reference's own CUDA kernels beside it (N = 1, rank 0; reported comparison) ----------------------------
    ref_cuda = None
    if rank == 0 and world == 1 and views == 1 and not args.no_reference_cuda:
        try:
            ref_cuda = time_reference_cuda(scene, H, W)
            if "ms_per_step" in ref_cuda:
                ref_cuda["ours_over_reference"] = ref_cuda["ms_per_step"] / ms_per_step
        except Exception as e:  # the comparison must never cost the bench line
            ref_cuda = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline (oracle port) on a bounded crop, rank 0 at N=1 only -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        oracle.set_num_threads(os.cpu_count() or 1)
        rows = args.cpu_rows or 256
        rows = min(H, max(16, rows // 16 * 16))
        t = cpu_step(scene_to_numpy(scene), rows)
        cpu = {"value": rows * W / t / 1e6, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
               "sample": f"top {rows} of {H} rows of the C4 frame ({W}x{rows} px, fwd+bwd, {t:.1f} s of wall time)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "gaussians": N, "sh_degree": 3, "texels_per_gaussian": 16,
                       "texture_channels": 3, "texture_layout": "(X,4) fp32: r, g, b at a 16-byte pitch" if rgba else "(X,3) fp32",
                       "width": W, "height": H, "views_per_step": views, "block_width": 16,
                       "settings": 256, "intersections_last_view": M, "max_intersections_seen": head["m_max"],
                       "l2": "inputs (0.9 GB of parameters, records and textures per view) are larger than the 126 MB L2",
                       "loss": head["loss"]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": head["launches"], "roofline": roofline, "cpu_baseline": cpu,
            "allreduce_ms": head["allreduce_ms"], "scale_base": scale_base, "reference_cuda": ref_cuda,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, scene, cams, mine, targets_host, dev, world, views):
    """The metric through the drop-in public API with host inputs: per step and per view, copy the camera matrices
    and the target image from pinned host memory, run project/AABB/count + texture_gaussians + loss + backward via
    autograd, and read the loss back."""
    import torch
    import torch.distributed as dist

    from gstex_cuda_b200 import sh as SH
    from gstex_cuda_b200.get_aabb_2d import get_aabb_2d, get_num_tiles_hit_2d, project_points
    from gstex_cuda_b200.loss import image_loss
    from gstex_cuda_b200.texture import texture_gaussians

    H, W, bw, intr = args.height, args.width, 16, scene["intrins"]
    leaves = {k: scene[k].clone().requires_grad_(True) for k in
              ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")}
    if args.texture_layout == "rgba":  # the same layout as the device-resident run: texels at a 16-byte pitch
        leaves["texture"] = torch.cat([scene["texture"], torch.zeros_like(scene["texture"][:, :1])], 1).contiguous().requires_grad_(True)
    # with several ranks the leaves' gradients are views of ONE flat buffer (autograd accumulates into them in place), so
    # that the data-parallel reduction is one collective, as in the fused step; a single rank lets autograd place them
    flat = None
    if world > 1:
        flat = torch.zeros(sum(t.numel() for t in leaves.values()), device=dev)
        off = 0
        for t in leaves.values():
            t.grad = flat[off:off + t.numel()].view_as(t)
            off += t.numel()
    cap = int(12 * args.points * max(1.0, args.scale_mult ** 2))
    cams_host = [(a.cpu().pin_memory(), b.cpu().pin_memory()) for a, b in cams]
    h2d = sum(targets_host[v].numel() * 4 + 2 * 64 for v in mine)

    # the step's inputs come from pinned host memory through the package's double-buffered loader: the copy of view k+1
    # runs on a copy stream underneath the kernels of view k.  Every timed step still copies every one of its views
    # (the copy for the first view of the next step is issued by the step before it; the one before the timed region
    # is issued by the warm-up, and the last timed step issues one for a step that never runs).
    from gstex_cuda_b200.prefetch import ViewPrefetcher
    loader = ViewPrefetcher(dev, depth=2)
    order = list(mine)

    def host_inputs(v):
        return (cams_host[v][0], cams_host[v][1], targets_host[v])

    loader.submit(host_inputs(order[0]))
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0, "pending": None}
    losses = []

    def read_pending():
        """Host value of the most recent step whose loss copy was issued (waits for that copy only)."""
        k = state["pending"]
        if k is None:
            return None
        loss_ready[k].synchronize()
        state["pending"] = None
        losses.append(float(loss_host[k][0]))
        return losses[-1]

    def one_step():
        total = None
        for idx, v in enumerate(order):
            vm, c2w, gt = loader.get()
            loader.submit(host_inputs(order[(idx + 1) % len(order)]))
            # with several views per step the two large gradients (texels, SH coefficients) are added straight into the
            # flat buffer by the backward kernels (the package's opt-in texture_grad= / coeffs_grad=), not returned to
            # autograd for a zero-fill + "grad += new" pass per view
            fused = dict(texture_grad=leaves["texture"].grad) if flat is not None else {}
            colors = SH.spherical_harmonics_colors(3, leaves["means"], c2w, leaves["sh_coeffs"],
                                                   coeffs_grad=leaves["sh_coeffs"].grad if flat is not None else None)
            _, depths = project_points(leaves["means"].detach(), vm, intr)
            centers, extents = get_aabb_2d(leaves["means"].detach(), leaves["scales"].detach(), 1.0,
                                           leaves["quats"].detach(), vm, intr)
            nth = get_num_tiles_hit_2d(centers, extents, H, W, bw)
            outs = texture_gaussians(scene["texture_info"], scene["texture_dims"], centers, extents, depths, nth, colors,
                                     leaves["opacities"], leaves["means"], leaves["scales"], 1.0, leaves["quats"],
                                     leaves["uv0"], leaves["umap"], leaves["vmap"], leaves["texture"], vm, c2w, *intr, H,
                                     W, bw, 1 << 8, scene["background"], max_intersects=cap, **fused)
            loss = image_loss(outs[4], outs[2], outs[5], gt)  # example.py:189-209, one kernel (csrc/loss.cu)
            loss.backward()
            loader.release()
            total = loss.detach() if total is None else total + loss.detach()
        if world > 1:
            dist.all_reduce(flat)
        # device -> host read of the step's result, every step: an asynchronous copy into pinned memory that is
        # consumed one step later, so that reading the loss of step k does not drain the queue before step k+1 is
        # enqueued (the usual way a training loop logs its loss); read_pending() collects the last one
        k = state["k"] & 1
        loss_host[k].copy_(total.reshape(1), non_blocking=True)
        loss_ready[k].record()
        val = read_pending()
        state["pending"], state["k"] = k, state["k"] + 1
        if flat is not None:
            flat.zero_()
        else:
            for t in leaves.values():
                t.grad = None
        return val

    # enough steps that one slow one (allocator growth, a host hiccup) does not decide the figure: up to 40 single-view
    # steps (0.2 s), up to 10 of the multi-view ones
    steps = max(3, min(args.steps, 40 if len(order) <= 2 else 10))
    loader_steps = [0]
    for _ in range(3):
        one_step()
    loader.bytes_copied = 0
    read_pending()  # the warm-up's last loss: nothing of the warm-up is left to read inside the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_read0 = len(losses)
    for _ in range(steps):
        one_step()
    read_pending()  # the last step's loss is read inside the timed region too
    e1.record()
    assert len(losses) - n_read0 == steps, "every timed step's loss must have been read back"
    loader_steps[0] = steps
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    return {"value": views * H * W / (ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 4, "ms_per_step": ms, "steps": steps,
            "api": "spherical_harmonics_colors + project_points + get_aabb_2d + get_num_tiles_hit_2d + texture_gaussians "
                   "(max_intersects capacity: no host sync per call"
                   + ("; texture_grad= / coeffs_grad=: texel and SH gradients added in place into the flat all-reduce buffer" if world > 1 else "")
                   + ") + "
                   "image_loss (the example.py loss), all autograd ops of the package; inputs staged by gstex_cuda_b200.prefetch.ViewPrefetcher "
                   "(pinned host -> device on a copy stream, one view ahead); the loss of every step is read back through "
                   "an asynchronous copy into pinned memory, consumed one step later",
            "h2d_bytes_measured_per_step": int(loader.bytes_copied // max(1, loader_steps[0]))}


if __name__ == "__main__":
    main()
