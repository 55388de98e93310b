"""GPU diagnostic for BASELINE config 1 (example.py --torch_compare): on IDENTICAL parameters, every iteration, the
reference's pure-PyTorch rasteriser (twin), the unmodified reference CUDA extension and this repo's kernels - all three
driven by the reference's own unmodified Python layer - are compared pairwise (max |d| per output and per gradient).
The parameters follow upstream's Adam trajectory on OUR gradients.  Prints one JSON line.
usage: python tools/c1_compare_diag.py [iterations]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_reference_python as R  # noqa: E402

sys.path.insert(0, R.REF)

PARAMS, OUTS, run = R.PARAMS, R.OUTS, R.three_way_c1


if __name__ == "__main__":
    rows = run(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
    for r in rows:
        worst = {n: max(r[n][k] for k in OUTS) for n in ("ours_vs_twin", "ref_vs_twin", "ours_vs_ref")}
        worst_g = {n: max(v for k, v in r[n].items() if k.startswith("grad_")) for n in ("ours_vs_twin", "ref_vs_twin", "ours_vs_ref")}
        print(f"it {r['iteration']:2d}  outputs max|d|: " + "  ".join(f"{n} {v:.2e}" for n, v in worst.items())
              + "   grads max|d|/max|g|: " + "  ".join(f"{n} {v:.2e}" for n, v in worst_g.items()), file=sys.stderr)
    print(json.dumps({"what": "C1 (example.py --torch_compare) three-way comparison on identical parameters; outputs: max |d|, "
                              "gradients: max |d| / max |g|", "rows": rows}))
