#!/bin/bash
# GPU visit for new components: run the given test files, then regenerate the reference-CUDA golden vectors.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_tests.sh "$@"
if [ -f oracle/_ref/gstex_ref_C.so ]; then
  timeout 600 python tests/golden/make_golden_ref_cuda.py > gpurun_out/make_golden.log 2>&1
  echo "== make_golden exit=$?" | tee -a gpurun_out/summary.txt
  tail -8 gpurun_out/make_golden.log
fi
