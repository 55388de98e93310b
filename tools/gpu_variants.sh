#!/bin/bash
# On the GPU box: bench every experiments/_variants/lib_*.so (kernel times only) against the production library.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cp gstex_cuda_b200/libgstex_b200.so /tmp/lib_prod.so
for so in /tmp/lib_prod.so experiments/_variants/lib_*.so; do
  cp $so gstex_cuda_b200/libgstex_b200.so
  n=$(basename $so .so)
  if [[ "$1" == test ]]; then timeout 300 python -m pytest tests/test_gpu_raster.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x 2>&1 | tail -1; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-scale-base --no-reference-cuda > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" <<'P'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/var_{n}.json'))
    print(f"{n:28s} step {d['ms_per_step']:.3f} ms  fwd {d['roofline']['kernel_ms']['raster_forward']:.3f}  bwd {d['roofline']['kernel_ms']['raster_backward']:.3f}")
except Exception as e:
    print(n, 'FAILED', e); print(open(f'gpurun_out/var_{n}.err').read()[-800:])
P
done
cp /tmp/lib_prod.so gstex_cuda_b200/libgstex_b200.so
