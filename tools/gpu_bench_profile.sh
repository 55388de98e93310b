#!/bin/bash
# bench (N=1) + ncu launch list + one ncu --set full capture of the two raster kernels.  Never a bench number under ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
echo "ref exit=$?"; cat gpurun_out/bench_ref_${TAG}.json
if [ "${2:-prof}" = "prof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:raster_ -s 2 -c 2 -f -o gpurun_out/prof_raster_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full exit=$?"; ls -la gpurun_out/*.ncu-rep
fi
