#!/bin/bash
# 2-GPU session: chain sharding == single GPU, pooled warmup over NCCL, bench scaling point
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== multi_gpu_check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/h_multi_gpu_check_$N.txt
echo "== bench --gpus $N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/h_bench_$N.err | tail -1 | tee gpurun_out/h_bench_$N.json | cut -c1-700
echo "== bench --gpus 1 (same box)"; timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/h_bench_1.json | cut -c1-400
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
