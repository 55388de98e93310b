"""1000 iterations of example.py's image overfit, the reference trainer (its own
UNMODIFIED Python - baseline/_ref - over its own CUDA extension - oracle/_ref -, torch.optim.Adam) next to this repo's
GStexTrainStep (one CUDA-graph launch per optimiser step), from the same initial parameters.  Reports the loss curves'
agreement and ms / iteration.  GPU box:   python tools/train_compare.py [C2|C3] [iterations]
"""
import importlib
import importlib.util
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_PY, REF_SO = os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref", "gstex_ref_C.so")
CONFIGS = {"C2": dict(height=256, width=256, num_points=100, num_texels=1000000),
           "C3": dict(height=256, width=256, num_points=10000, num_texels=0)}


def load_reference():
    spec = importlib.util.spec_from_file_location("gstex_ref_C", REF_SO)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    sys.path.insert(0, REF_PY)
    pkg = importlib.import_module("gstex_cuda")
    sys.modules["gstex_cuda.cuda"] = ext  # what gstex_cuda/cuda/_backend.py resolves to upstream
    pkg.cuda = ext
    return importlib.import_module("example")


def gt_image(h, w):
    gt = torch.ones((h, w, 3))
    gt[: h // 2, : w // 2, :] = torch.tensor([1.0, 0.0, 0.0])
    gt[h // 2:, w // 2:, :] = torch.tensor([0.0, 0.0, 1.0])
    return gt


def run(name, iterations):
    from gstex_cuda_b200.trainer import GStexTrainStep

    cfg = CONFIGS[name]
    example = load_reference()
    example.seed_everything(1)
    tr = example.SimpleTrainer(gt_image=gt_image(cfg["height"], cfg["width"]), num_points=cfg["num_points"],
                               num_texels=cfg["num_texels"])
    H, W, N = tr.H, tr.W, tr.num_points
    raw = dict(means=tr.means.detach().clone(), scales=tr.scales.detach().clone(), quats=tr.quats.detach().clone(),
               opacities=tr.opacities.detach().clone(), rgbs=tr.rgbs.detach().clone(), texture=tr.texture.detach().clone(),
               mapping=tr.mapping.detach().clone())
    dims = torch.zeros(N, 3, dtype=torch.int32, device="cuda:0")
    dims[:, 0], dims[:, 1] = tr.th, tr.tw
    dims[:, 2] = torch.arange(N, dtype=torch.int32, device="cuda:0") * (tr.th * tr.tw)
    intr = (tr.focal, tr.focal, W / 2, H / 2)

    # ---- ours: one graph launch per optimiser step, losses kept on the device
    ours = GStexTrainStep(raw, dims, H, W, intrins=intr, sh_degree=3, lr=1e-2, background=tr.background.clone())
    cams = [(tr.viewmat.clone().contiguous(), tr.c2w.clone().contiguous())]
    targets = [tr.gt_image.clone().contiguous()]
    ours.capture(cams, targets)
    losses_o = torch.zeros(iterations, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(iterations):
        losses_o[it:it + 1].copy_(ours.replay(), non_blocking=True)
    torch.cuda.synchronize()
    t_ours = time.perf_counter() - t0
    ours.fused.check_overflow()

    # ---- the reference trainer, as upstream runs it (prints a loss per iteration: that .item() is part of its loop)
    import contextlib
    import io
    buf = io.StringIO()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(buf):
        tr.train(iterations=iterations, lr=1e-2, save_imgs=False, torch_compare=False)
    torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0
    losses_r = [float(l.split("Loss:")[1]) for l in buf.getvalue().splitlines() if "Loss:" in l]
    # control: the reference trainer AGAIN from the same seed - its backward sums with atomics in arbitrary order, so two
    # of its own runs drift apart as well; that drift is the yardstick for "the curves agree"
    example.seed_everything(1)
    tr2 = example.SimpleTrainer(gt_image=gt_image(cfg["height"], cfg["width"]), num_points=cfg["num_points"],
                                num_texels=cfg["num_texels"])
    buf2 = io.StringIO()
    with contextlib.redirect_stdout(buf2):
        tr2.train(iterations=iterations, lr=1e-2, save_imgs=False, torch_compare=False)
    losses_r2 = [float(l.split("Loss:")[1]) for l in buf2.getvalue().splitlines() if "Loss:" in l]
    rel_rr = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(losses_r2, losses_r)]
    lo = losses_o.cpu().tolist()
    rel = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(lo, losses_r)]
    marks = [i for i in (0, 9, 99, 499, iterations - 1) if i < iterations]
    out = {"config": name, **cfg, "iterations": iterations, "texels_per_gaussian": tr.th * tr.tw,
           "ms_per_iter": {"ours_graph_replay": 1e3 * t_ours / iterations, "reference_trainer": 1e3 * t_ref / iterations,
                           "speedup": t_ref / t_ours},
           "loss": {"ours": {str(i + 1): lo[i] for i in marks}, "reference": {str(i + 1): losses_r[i] for i in marks},
                    "relative_difference": {str(i + 1): rel[i] for i in marks},
                    "max_relative_difference_first_100": max(rel[:100]), "max_relative_difference": max(rel),
                    "final_ratio_ours_over_reference": lo[-1] / losses_r[-1],
                    "reference_vs_reference_rerun": {"relative_difference": {str(i + 1): rel_rr[i] for i in marks},
                                                     "max_relative_difference_first_100": max(rel_rr[:100]),
                                                     "max_relative_difference": max(rel_rr),
                                                     "final_ratio": losses_r2[-1] / losses_r[-1]}},
           "max_intersections_seen": int(ours.fused.max_count_seen.item()),
           "note": "same initial parameters, same Adam hyper-parameters; the two runs diverge slowly through fp32 "
                   "summation order (atomics) amplified by 1000 Adam steps"}
    print(json.dumps(out))


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "C2", int(sys.argv[2]) if len(sys.argv) > 2 else 1000)
