"""Host-side cost of one step of bench.py's e2e path (the drop-in autograd API): the same calls on a scene so small that
the GPU time is negligible, so that ms/step IS the CPU time spent enqueueing a step (Python, autograd, ctypes, launches).
If that approaches the GPU time of a C4 step (4.2 ms) the e2e figure becomes host-bound on a slow or busy host."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gstex_cuda_b200.scenes import synthetic_scene

dev = torch.device("cuda:0")
args = types.SimpleNamespace(height=64, width=64, points=2000, steps=10, texture_layout="rgba", scale_mult=1.0)
scene = synthetic_scene(args.points, args.width, args.height, seed=1, device=dev)
cams = [(scene["viewmat"], scene["c2w"])]
targets_host = {0: torch.rand(args.height, args.width, 3).pin_memory()}
for _ in range(3):
    r = bench.run_e2e(args, scene, cams, [0], targets_host, dev, 1, 1)
    print("ms per step on a 2000-Gaussian 64x64 scene (host-bound):", round(r["ms_per_step"], 3))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
bench.run_e2e(args, scene, cams, [0], targets_host, dev, 1, 1)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
