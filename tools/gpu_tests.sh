#!/bin/bash
# Run every -m gpu test file in its own process (a faulting kernel must not poison the others) and keep the logs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
rc_all=0
for f in ${@:-tests/test_gpu_reference_python.py tests/test_gpu_sh_sample.py tests/test_gpu_binning.py tests/test_gpu_raster.py tests/test_gpu_api.py tests/test_gpu_pipeline.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_texture_edit.py tests/test_gpu_train_ops.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py}; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -s --timeout 600 > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $f exit=$rc" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed|error" gpurun_out/$name.log | tail -2 | tee -a gpurun_out/summary.txt
  [ $rc -ne 0 ] && rc_all=1
done
grep -E "^(FAILED|ERROR)" gpurun_out/*.log | head -60
exit $rc_all
