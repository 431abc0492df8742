#!/bin/bash
# multi-GPU session (run with gpurun --gpus N): the driver's own launch line for bench.py, the D2H ceiling probe, the NCCL checks
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
N=${1:-2}
echo "== d2h ceiling, $N ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/r2/d2h_multi.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -$((N+3))
echo "== multi_gpu_check, $N ranks"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/multi_gpu_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -8
echo "== bench --gpus $N"; NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1; tail -1 gpurun_out/r2_bench_${N}gpu.log | cut -c1-3500
echo "== bench --impl reference --gpus $N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-600
