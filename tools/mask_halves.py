"""What a half-warp-decoupled walk would save on the C4 view: per (tile, warp) count the entries whose blend mask touches
the left / right 4x4 half of the 8x4 patch (lane = y*8+x) and compare max(nL, nR) with the union."""
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
m = fused.masks[:M].to(torch.int64) & 0xffffffff          # (M, 8 warps)
bins = fused.tile_bins.to(torch.int64)
tile_of = torch.repeat_interleave(torch.arange(bins.shape[0], device=dev), (bins[:, 1] - bins[:, 0]))
assert tile_of.numel() == M
def per_tile_warp(flag):  # flag (M, 8) bool -> (tiles, 8) counts
    out = torch.zeros((bins.shape[0], 8), dtype=torch.int64, device=dev)
    out.index_add_(0, tile_of, flag.to(torch.int64))
    return out
for name, lm, rm in (("4x4 halves", 0x0f0f0f0f, 0xf0f0f0f0), ("8x2 halves", 0x0000ffff, 0xffff0000)):
    nU = per_tile_warp(m != 0); nL = per_tile_warp((m & lm) != 0); nR = per_tile_warp((m & rm) != 0)
    mx = torch.maximum(nL, nR)
    print(f"{name}: blended (warp,entry) union {int(nU.sum())}  left {int(nL.sum())}  right {int(nR.sum())}  "
          f"sum max(L,R) {int(mx.sum())}  ratio {float(mx.sum()) / float(nU.sum()):.3f}")
