#!/bin/bash
# Round-2 measurement set on one GPU: full -m gpu suite, bench (all legs), CPU reference arm, ncu launch list + full capture,
# C2 / C3 1000-iteration training comparison.   Usage: tools/gpu_r2_final.sh <tag>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
tag=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "== smoke exit=$?" | tee -a gpurun_out/summary.txt
bash tools/gpu_tests.sh tests/test_gpu_reference_python.py tests/test_gpu_sh_sample.py tests/test_gpu_binning.py tests/test_gpu_raster.py tests/test_gpu_api.py tests/test_gpu_pipeline.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_texture_edit.py tests/test_gpu_train_ops.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py tests/test_bench_contract.py
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "== bench exit=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "== bench reference exit=$?" | tee -a gpurun_out/summary.txt
for c in C2 C3; do
  timeout 900 python tools/train_compare.py $c 1000 > gpurun_out/train_compare_${c}_$tag.json 2> gpurun_out/train_compare_${c}_$tag.err
  echo "== train_compare $c exit=$?" | tee -a gpurun_out/summary.txt
  tail -c 1500 gpurun_out/train_compare_${c}_$tag.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-scale-base --no-reference-cuda \
    > gpurun_out/ncu_launch_$tag.log 2>&1
echo "== ncu launches exit=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 2 -c 2 \
    -f -o gpurun_out/prof_raster_$tag python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-scale-base --no-reference-cuda \
    > gpurun_out/ncu_full_$tag.log 2>&1
echo "== ncu full exit=$?" | tee -a gpurun_out/summary.txt
python - "$tag" <<'P'
import json,sys
d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
print("bench: step %.3f ms  value %.1f  e2e %s  kernels %s" % (d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), {k: round(v,3) for k,v in d['roofline']['kernel_ms'].items()}))
print("ref_cuda", d['reference_cuda'].get('ms_per_step'), "scale_base", d['scale_base']['value'], "cpu", d['cpu_baseline'])
P
cat gpurun_out/summary.txt
