#!/bin/bash
# Multi-GPU visit: data-parallel correctness check + bench at N ranks (NCCL logs to stderr) + the CPU reference arm under
# torchrun.   Usage: tools/gpu_multi.sh <N> <tag>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1; tag=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_multi_$tag.txt 2>&1
timeout 600 $TR --master-port 29533 tools/check_dp.py > gpurun_out/check_dp_${tag}.json 2> gpurun_out/check_dp_${tag}.err
echo "== check_dp exit=$?"; cat gpurun_out/check_dp_${tag}.json
NCCL_DEBUG=INFO timeout 900 $TR --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
echo "== bench N=$N exit=$?  stdout lines: $(wc -l < gpurun_out/bench_n${N}_$tag.json)"
grep -c "NCCL INFO" gpurun_out/bench_n${N}_$tag.err | sed 's/^/NCCL INFO lines on stderr: /'
grep -m3 -E "NVLS|nranks|Connected all" gpurun_out/bench_n${N}_$tag.err | cut -c1-200
python - "$N" "$tag" <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}_{sys.argv[2]}.json').read().strip().splitlines()[-1])
print("bench N=%s: step %.2f ms  value %.1f  e2e %.1f  allreduce_ms %s  launches %d" % (sys.argv[1], d['ms_per_step'], d['value'], d['e2e']['value'], d['allreduce_ms'], d['gpu_launches']))
P
timeout 600 $TR --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n${N}_$tag.json 2> gpurun_out/bench_ref_n${N}_$tag.err
echo "== bench reference N=$N exit=$?  stdout lines: $(wc -l < gpurun_out/bench_ref_n${N}_$tag.json)"
tail -c 400 gpurun_out/bench_n${N}_$tag.err | head -5
