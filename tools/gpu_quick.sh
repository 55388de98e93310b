#!/bin/bash
# quick GPU iteration: raster parity tests + bench (no CPU baseline) [+ ncu full of the raster kernels with "ncu"]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
bash tools/gpu_tests.sh tests/test_gpu_raster.py tests/test_gpu_api.py tests/test_gpu_vs_reference_cuda.py
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'P'
import json
try:
    d=json.load(open('gpurun_out/bench_n1.json'))
    print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], d['roofline']['kernel_ms'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench_n1.err').read()[-2000:])
P
if [[ "$1" == ncu ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 2 -c 2 \
      -f -o gpurun_out/prof_raster python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu exit=$?"
fi
