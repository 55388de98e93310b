"""GPU diagnostic: which fp32 association reproduces torch's `points @ viewmat.T[:3,:3] + t` depth bit for bit
(the reference's get_aabb_2d.py:22-32), and does our project_points match it?  Prints mismatch counts per candidate."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gstex_cuda_b200.scenes import synthetic_scene, arc_cameras
from gstex_cuda_b200 import get_aabb_2d as A

dev = "cuda:0"
s = synthetic_scene(1_000_000, 1920, 1080, seed=1234, device=dev)
P = s["means"]
def fma(a, b, c):  # exact fp32 fma through fp64 (24x24-bit products are exact in fp64; one rounding to fp64, one to fp32)
    return (a.double() * b.double() + c.double()).float()
for name, vm in [("front", s["viewmat"])] + [(f"arc{i}", arc_cameras(64)[i][0].to(dev)) for i in (0, 45)]:
    vp = P @ vm.T[:3, :3] + vm.T[3:, :3]
    ref = vp[:, 2].contiguous()
    m = [vm[2, k].item() for k in range(4)]
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    comp = {"x": (x, m[0]), "y": (y, m[1]), "z": (z, m[2])}
    res = {}
    for order in itertools.permutations("xyz"):
        a, b, c = [comp[o] for o in order]
        acc = a[0] * a[1]
        acc = fma(b[0], torch.full_like(x, b[1]), acc)
        acc = fma(c[0], torch.full_like(x, c[1]), acc)
        res["fma chain " + "".join(order) + " then +t"] = int((acc + m[3] != ref).sum())
        acc2 = fma(a[0], torch.full_like(x, a[1]), torch.full_like(x, m[3]))
        acc2 = fma(b[0], torch.full_like(x, b[1]), acc2)
        acc2 = fma(c[0], torch.full_like(x, c[1]), acc2)
        res["fma chain from t " + "".join(order)] = int((acc2 != ref).sum())
    res["plain mul/add xyz"] = int((((x * m[0] + y * m[1]) + z * m[2]) + m[3] != ref).sum())
    _, ours = A.project_points(P, vm.contiguous(), s["intrins"])
    res["OURS project_points"] = int((ours != ref).sum())
    print(name, {k: v for k, v in sorted(res.items(), key=lambda kv: kv[1])[:5]}, "| ours:", res["OURS project_points"])
