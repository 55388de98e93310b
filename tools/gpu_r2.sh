#!/bin/bash
# One GPU-box visit of round 2.  Usage: tools/gpu_r2.sh <tag> [tests] [bench] [variants] [ncu] [diag] [ref]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
tag=$1; shift
has() { for a in "$@"; do [[ " $STAGES " == *" $a "* ]] && return 0; done; return 1; }
STAGES="$*"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/smi.txt 2>&1
if has tests; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "== smoke exit=$?" | tee -a gpurun_out/summary.txt
  bash tools/gpu_tests.sh ${TEST_FILES}
fi
if has diag; then
  timeout 300 python tools/depth_assoc.py > gpurun_out/depth_assoc_$tag.txt 2>&1; cat gpurun_out/depth_assoc_$tag.txt
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "== bench exit=$?" | tee -a gpurun_out/summary.txt
  python - "$tag" <<'P'
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
    print("bench: step %.3f ms  value %.1f  e2e %s  kernels %s" % (d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()}))
    print("clocks", d['clocks'], "M", d['config']['intersections_last_view'])
except Exception as e:
    print('bench FAILED', e); print(open(f'gpurun_out/bench_{sys.argv[1]}.err').read()[-1500:])
P
fi
if has variants; then bash tools/gpu_variants.sh; fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-scale-base --no-reference-cuda \
      > gpurun_out/ncu_launch_$tag.log 2>&1
  echo "== ncu launches exit=$?" | tee -a gpurun_out/summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 2 -c 2 \
      -f -o gpurun_out/prof_raster_$tag python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-scale-base --no-reference-cuda \
      > gpurun_out/ncu_full_$tag.log 2>&1
  echo "== ncu full exit=$?" | tee -a gpurun_out/summary.txt
fi
if has ref && [ -f oracle/_ref/gstex_ref_C.so ]; then
  timeout 600 python tools/time_reference_cuda.py > gpurun_out/ref_cuda_time_$tag.json 2> gpurun_out/ref_cuda_time_$tag.err
  echo "== reference CUDA timing exit=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/ref_cuda_time_$tag.json
fi
cat gpurun_out/summary.txt
