#!/bin/bash
# One GPU-box visit: parity tests, golden vectors from the reference CUDA extension, bench (ours + CPU arm),
# the reference CUDA rasteriser timed beside it, the ncu launch list and one ncu --set full capture of the
# two rasteriser kernels.  Everything lands in gpurun_out/.   Usage: tools/gpu_round.sh [tests|bench|ncu|all]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt

if [[ $what == all || $what == tests ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "== smoke exit=$?" | tee -a gpurun_out/summary.txt
  bash tools/gpu_tests.sh
  if [ -f oracle/_ref/gstex_ref_C.so ]; then
    timeout 600 python tests/golden/make_golden_ref_cuda.py > gpurun_out/make_golden.log 2>&1
    echo "== make_golden exit=$?" | tee -a gpurun_out/summary.txt
  fi
fi

if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  echo "== bench exit=$?" | tee -a gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench_n1.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "== bench reference exit=$?" | tee -a gpurun_out/summary.txt
  if [ -f oracle/_ref/gstex_ref_C.so ]; then
    timeout 600 python tools/time_reference_cuda.py > gpurun_out/ref_cuda_time.json 2> gpurun_out/ref_cuda_time.err
    echo "== reference CUDA timing exit=$?" | tee -a gpurun_out/summary.txt
    cat gpurun_out/ref_cuda_time.json
  fi
fi

if [[ $what == all || $what == ncu ]]; then
  # launch list of the bench command (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline \
      > gpurun_out/ncu_launches.log 2>&1
  echo "== ncu launches exit=$?" | tee -a gpurun_out/summary.txt
  # full capture of the two rasteriser kernels (second view = warm)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 2 -c 2 \
      -f -o gpurun_out/prof_raster python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline \
      > gpurun_out/ncu_full.log 2>&1
  echo "== ncu full exit=$?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
