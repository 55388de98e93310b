"""Kernel timeline of one fused C4 step (FusedTrainStep.step, one view): start / end of every kernel relative to the
step's first kernel, with its stream - shows what overlaps what (torch.profiler; not a timing run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gstex_cuda_b200.pipeline import FusedTrainStep
from gstex_cuda_b200.scenes import synthetic_scene

dev = torch.device("cuda:0")
H, W, N = 1080, 1920, 1_000_000
scene = synthetic_scene(N, W, H, seed=1234, device=dev)
scene["texture"] = torch.cat([scene["texture"], torch.zeros_like(scene["texture"][:, :1])], 1).contiguous()
params = {k: scene[k] for k in ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")}
st = FusedTrainStep(params, scene["texture_dims"], H, W, intrins=scene["intrins"], background=scene["background"],
                    sh_degree=3, max_intersects=12 * N, texture_rgba=True)
cams, tg = [(scene["viewmat"], scene["c2w"])], [torch.rand(H, W, 3, device=dev)]
for _ in range(5):
    st.step(cams, tg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st.step(cams, tg)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(ev) if "sh_forward" in e.name or "project_aabb" in e.name]
# the middle step: from its first head kernel to the next step's
firsts = [i for k, i in enumerate(starts) if k == 0 or ev[i].time_range.start - ev[starts[k - 1]].time_range.start > 1000]
i0, i1 = firsts[1], firsts[2]
t0 = ev[i0].time_range.start
for e in ev[i0:i1]:
    print(f"{(e.time_range.start - t0) / 1e3:8.3f} -> {(e.time_range.end - t0) / 1e3:8.3f} ms  {e.name[:70]}")
print(f"step span {(ev[i1].time_range.start - t0) / 1e3:.3f} ms")
