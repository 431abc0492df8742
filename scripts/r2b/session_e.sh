#!/bin/bash
# round 2b, 2 GPUs (gpurun --gpus 2): the NCCL checks and the driver's own launch line for bench.py with the final kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
N=2
echo "== multi_gpu_check, $N ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -8 | cut -c1-400
echo "== bench --gpus $N"; NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_final_bench_${N}gpu.log 2>&1; tail -1 gpurun_out/r2_final_bench_${N}gpu.log | cut -c1-3800
