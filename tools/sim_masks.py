"""CPU simulation of the forward pass's blend masks on a crop of the C4 view (float64 affine-form model), used to size
the forward / backward work queues without a GPU.   python tools/sim_masks.py [rows] [row0]"""
import os, sys, math, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from gstex_cuda_b200.scenes import synthetic_scene

K_SIGMA = math.sqrt(0.5 * math.log2(math.e))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, W, Hfull = 1_000_000, 1920, 1080
sc = synthetic_scene(N, W, Hfull, seed=1234)
s = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in sc.items()}
fx, fy, cx, cy = s["intrins"]
H = rows
b = oracle.bin_view(s["means"], s["scales"], 1.0, s["quats"], s["viewmat"], s["intrins"], H, W, 16)
ids, bins = b["gaussian_ids_sorted"], b["tile_bins"]
print("M", len(ids), "tiles", bins.shape[0])
used = np.unique(ids)
# vectorised pack (tests/formulation.py::pack_record), identity-rotation camera at z=-8 handled generally
c2w = s["c2w"].astype(np.float64); Rc = c2w[:3, :3]; o = c2w[:3, 3]
q = s["quats"][used].astype(np.float64); w_, x, y, z = q.T
a1 = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w_ * z), 2 * (x * z - w_ * y)], -1)
a2 = np.stack([2 * (x * y - w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w_ * x)], -1)
a3 = np.stack([2 * (x * z + w_ * y), 2 * (y * z - w_ * x), 1 - 2 * (x * x + y * y)], -1)
d = s["means"][used].astype(np.float64) - o
c0 = (a3 * d).sum(-1); b1 = (a1 * d).sum(-1); b2 = (a2 * d).sum(-1)
w1 = c0[:, None] * a1 - b1[:, None] * a3; w2 = c0[:, None] * a2 - b2[:, None] * a3
h1 = w1 @ Rc; h2 = w2 @ Rc; h3 = a3 @ Rc
k1 = K_SIGMA / s["scales"][used, 0].astype(np.float64); k2 = K_SIGMA / s["scales"][used, 1].astype(np.float64)
mc = d @ Rc
rc = np.stack([mc[:, 0] / mc[:, 2], mc[:, 1] / mc[:, 2], np.ones(len(used))], -1)
rec = dict(xc=fx * rc[:, 0] + cx, yc=fy * rc[:, 1] + cy, c0=c0, opac=s["opacities"][used, 0].astype(np.float64),
           P1=np.stack([k1 * h1[:, 0] / fx, k1 * h1[:, 1] / fy], -1), P2=np.stack([k2 * h2[:, 0] / fx, k2 * h2[:, 1] / fy], -1),
           A3=np.stack([h3[:, 0] / fx, h3[:, 1] / fy], -1), c3=(h3 * rc).sum(-1))
lut = np.full(N, -1, np.int64); lut[used] = np.arange(len(used))
tiles_x = W // 16
# pixel layout: thread tr -> warp w = tr>>5, lane l: lx = (w&1)*8 + (l&7), ly = (w>>1)*4 + (l>>3)
tr = np.arange(256); wv = tr >> 5; l = tr & 31
lx = ((wv & 1) << 3) + (l & 7); ly = ((wv >> 1) << 2) + (l >> 3)
all_masks = []  # per tile: (G, 8) uint32 blend masks ; alive (G, 8) bool = warp not finished before the entry
t0 = time.time()
for t in range(bins.shape[0]):
    lo, hi = bins[t]
    if hi <= lo: all_masks.append((np.zeros((0, 8), np.uint32), np.zeros((0, 8), bool))); continue
    g = lut[ids[lo:hi]]
    tx, ty = t % tiles_x, t // tiles_x
    px = tx * 16 + lx + 0.5; py = ty * 16 + ly + 0.5
    u = (px - cx) / fx; v = (py - cy) / fy
    rn = np.sqrt(u * u + v * v + 1.0)
    ex = px[None, :] - rec["xc"][g][:, None]; ey = py[None, :] - rec["yc"][g][:, None]
    n1 = rec["P1"][g][:, :1] * ex + rec["P1"][g][:, 1:] * ey
    n2 = rec["P2"][g][:, :1] * ex + rec["P2"][g][:, 1:] * ey
    D = rec["A3"][g][:, :1] * ex + rec["A3"][g][:, 1:] * ey + rec["c3"][g][:, None]
    D = np.where(np.abs(D) < 1e-6, 1e-6, D)
    qq = (n1 * n1 + n2 * n2) / (D * D)
    alpha = np.minimum(0.99, rec["opac"][g][:, None] * np.exp2(-qq))
    tt = rec["c0"][g][:, None] / D * rn[None, :]
    skipped = (tt < 0.01) | (tt > 1000) | (alpha < 1 / 255)
    aeff = np.where(skipped, 0.0, alpha)
    Tb = np.cumprod(np.vstack([np.ones((1, 256)), 1 - aeff[:-1]]), 0)  # T before each entry (ignoring stop)
    stop = Tb * (1 - alpha) <= 1e-4
    done = np.maximum.accumulate(stop, 0)
    blend = (~skipped) & (~done)
    bl = blend.reshape(-1, 8, 32)
    masks = (bl * (1 << np.arange(32, dtype=np.uint64))[None, None, :]).sum(-1).astype(np.uint32)
    dn = done.reshape(-1, 8, 32).all(-1)
    alive = ~np.vstack([np.zeros((1, 8), bool), dn[:-1]])
    all_masks.append((masks, alive))
print("sim time %.1f s" % (time.time() - t0))
np.savez_compressed("/tmp/sim_masks_%d.npz" % rows, **{"m%d" % i: m for i, (m, a) in enumerate(all_masks)},
                    **{"a%d" % i: a for i, (m, a) in enumerate(all_masks)})
pc = np.concatenate([np.array([bin(int(v)).count("1") for v in m.reshape(-1)]) for m, a in all_masks])
nz = pc[pc > 0]
print("warp-entries", pc.size, "non-empty", nz.size, "pairs", nz.sum(), "mean lanes %.2f" % nz.mean())
