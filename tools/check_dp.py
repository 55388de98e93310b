"""Multi-GPU correctness check (run under torchrun on N GPUs): the view-sharded step + ONE NCCL all-reduce of the gradient
arena must equal the single-process step over the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_dp.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG"):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
import torch
import torch.distributed as dist
from gstex_cuda_b200.pipeline import FusedTrainStep, DataParallelTrainStep
from gstex_cuda_b200.scenes import synthetic_scene, arc_cameras

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, W, H, V = 200_000, 960, 544, 8
s = synthetic_scene(N, W, H, seed=11, device=dev)
params = {k: s[k] for k in ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")}
# the benchmark's layout: texels at a 16-byte pitch, read and differentiated in place
params["texture"] = torch.cat([s["texture"], torch.zeros_like(s["texture"][:, :1])], 1).contiguous()
cams = [(a.to(dev), b.to(dev)) for a, b in arc_cameras(V)]
g = torch.Generator().manual_seed(3)
targets = [torch.rand(H, W, 3, generator=g).to(dev) for _ in range(V)]
mk = lambda: FusedTrainStep(params, s["texture_dims"], H, W, intrins=s["intrins"], sh_degree=3, background=s["background"],
                            max_intersects=16 * N, texture_rgba=True)
dp = DataParallelTrainStep(mk(), rank, world)
loss_dp = dp.step(cams, targets)
torch.cuda.synchronize()
if rank == 0:
    ref = mk()
    loss_ref = ref.step(cams, targets)
    torch.cuda.synchronize()
    a, b = dp.inner.grad_arena.double(), ref.grad_arena.double()
    scale = float(b.abs().max())
    bad = ((a - b).abs() > 1e-5 * scale + 1e-3 * b.abs()).double().mean()
    out = {"world_size": world, "views": V, "gaussians": N, "arena_floats": a.numel(), "loss_dp": float(loss_dp),
           "loss_single": float(loss_ref), "max_abs_diff_over_max": float((a - b).abs().max()) / scale,
           "fraction_out_of_tolerance": float(bad), "tolerance": "1e-5 * max|g| + 1e-3 * |g| (atomic-order noise)"}
    print(json.dumps(out))
    assert abs(out["loss_dp"] - out["loss_single"]) <= 1e-5 * abs(out["loss_single"]) and float(bad) < 1e-4
dist.barrier()
dist.destroy_process_group()
