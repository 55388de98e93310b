"""Small end-to-end run of every hot-path kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gstex_cuda_b200.scenes import random_small_scene, synthetic_scene
from gstex_cuda_b200.pipeline import FusedTrainStep
from gpu_util import bin_cuda, forward_cuda, backward_cuda, random_vout

for settings, C in ((1 << 8, 3), ((1 << 8) | (1 << 9) | (1 << 10), 3), (1 << 2, 5), ((1 << 8) | (1 << 16) | (24 << 17) | (4 << 26), 3)):
    s = random_small_scene(400, 96, 80, seed=3, channels=C, device="cuda:0")
    s["settings"] = settings
    b = bin_cuda(s)
    f, scratch = forward_cuda(s, b["gaussian_ids_sorted"], b["tile_bins"])
    if not settings & (1 << 16):
        backward_cuda(s, b["gaussian_ids_sorted"], b["tile_bins"], f, random_vout(s, 0), scratch=scratch)
s = synthetic_scene(20000, 320, 192, seed=7, device="cuda:0")
fused = FusedTrainStep({k: s[k] for k in ("means", "scales", "quats", "opacities", "sh_coeffs", "uv0", "umap", "vmap", "texture")},
                       s["texture_dims"], s["H"], s["W"], intrins=s["intrins"], sh_degree=3, background=s["background"],
                       max_intersects=40 * s["num_points"])
cams = [(s["viewmat"], s["c2w"])] * 2
fused.step(cams, [s["target"], s["target"]])
torch.cuda.synchronize()
print("sanitize smoke done, M =", fused.check_overflow())
