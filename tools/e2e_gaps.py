"""GPU idle gaps inside one e2e step (drop-in API path of bench.py): torch.profiler kernel timeline -> the gaps > 5 us
and the kernel that follows each."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gstex_cuda_b200.scenes import synthetic_scene
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
args = types.SimpleNamespace(height=1080, width=1920, points=1_000_000, steps=3, texture_layout="rgba", scale_mult=1.0)
scene = synthetic_scene(args.points, args.width, args.height, seed=1234, device=dev)
cams = [(scene["viewmat"], scene["c2w"])]
targets_host = {0: torch.rand(args.height, args.width, 3).pin_memory()}
# run_e2e times `steps` steps after 2 warm-ups; profile the whole call and analyse the last step
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    r = bench.run_e2e(args, scene, cams, [0], targets_host, dev, 1, 1)
print("e2e ms/step", r["ms_per_step"])
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last step = from the last sh_forward kernel
starts = [i for i, e in enumerate(ev) if "sh_forward" in e.name]
# a MIDDLE step (second to last): the last one ends with the read-back of the final loss and the closing synchronise
i0, i1 = starts[-2], starts[-1]
step = ev[i0:i1 + 1]
t_end = step[0].time_range.start
busy = 0.0
gaps = []
prev = ""
for e in step:
    s, t = e.time_range.start, e.time_range.end
    if s > t_end + 5:
        gaps.append((s - t_end, f"{e.name[:48]:48s} (after {prev[:40]}; at +{(s - step[0].time_range.start)/1e3:.3f} ms)"))
    prev = e.name
    busy += max(0, t - max(s, t_end))
    t_end = max(t_end, t)
span = t_end - step[0].time_range.start
print(f"second-to-last step (first kernel to the next step's first kernel): span {span/1e3:.3f} ms, busy {busy/1e3:.3f} ms, idle {(span-busy)/1e3:.3f} ms in {len(gaps)} gaps > 5 us")
for g, n in sorted(gaps, reverse=True)[:15]:
    print(f"  {g:8.1f} us before {n}")
# per-kernel totals of the last step (what the API path runs beyond the fused step's kernels)
tot = {}
for e in step:
    d = e.time_range.end - e.time_range.start
    k = e.name[:70]
    tot[k] = (tot.get(k, (0, 0))[0] + d, tot.get(k, (0, 0))[1] + 1)
print("kernels of the last step (total us, count):")
for k, (d, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {d:9.1f} us  x{c:<3d} {k}")
