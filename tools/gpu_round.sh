#!/bin/bash
# one GPU call: all -m gpu tests, reference-CUDA golden dump + timing, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
rm -f gpurun_out/summary.txt
bash tools/gpu_tests.sh tests/test_gpu_sh_sample.py tests/test_gpu_binning.py tests/test_gpu_raster.py tests/test_gpu_api.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_pipeline.py
timeout 300 python tests/golden/make_golden_ref_cuda.py > gpurun_out/golden_ref.log 2>&1; tail -6 gpurun_out/golden_ref.log
timeout 600 python tools/time_reference_cuda.py > gpurun_out/reference_cuda_timing.json 2> gpurun_out/reference_cuda_timing.err; cat gpurun_out/reference_cuda_timing.json; tail -3 gpurun_out/reference_cuda_timing.err
bash tools/gpu_bench_profile.sh ${1:-r01b} noprof
