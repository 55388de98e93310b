"""Distribution of the forward pass's blend masks on the C4 view: how many pixels of a warp's 8x4 patch blend an entry."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gstex_cuda_b200.pipeline import FusedTrainStep
from gstex_cuda_b200.scenes import synthetic_scene
dev = torch.device('cuda:0')
H, W, N = 1080, 1920, 1000000
scene = synthetic_scene(N, W, H, seed=1234, device=dev)
params = {k: scene[k] for k in ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")}
fused = FusedTrainStep(params, scene["texture_dims"], H, W, intrins=scene["intrins"], sh_degree=3, background=scene["background"], max_intersects=12 * N)
fused.begin_step(); fused.view_forward(scene["viewmat"], scene["c2w"]); torch.cuda.synchronize()
M = int(fused.num_isect.item())
m = fused.masks[:M].reshape(-1).to(torch.int64) & 0xffffffff
pc = torch.zeros_like(m)
for b in range(32): pc += (m >> b) & 1
nz = pc[pc > 0]
hist = torch.bincount(nz, minlength=33).double()
pairs = hist * torch.arange(33, dtype=torch.double, device=dev)
print("M", M, "warp-entries", m.numel(), "non-empty", nz.numel(), "pairs", int(pairs.sum()), "mean lanes", float(pairs.sum() / nz.numel()))
ce = torch.cumsum(hist, 0) / hist.sum(); cp = torch.cumsum(pairs, 0) / pairs.sum()
for k in (1, 2, 4, 8, 12, 16, 20, 24, 28, 31, 32):
    print(f"popc<={k:2d}: entries {100*float(ce[k]):5.1f}%  pairs {100*float(cp[k]):5.1f}%")
