#!/bin/bash
# 8 GPUs (charged 8x): the host-side D2H ceiling and the driver's own launch line for bench.py at N = 8
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
N=8
echo "== d2h ceiling, $N ranks"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/r2/d2h_multi.py 2>&1 | grep -v "^W\|^\[W\|warn\|^\*\*\*\|OMP_NUM" | tail -$((N+3))
echo "== bench --gpus $N"; NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1; tail -1 gpurun_out/r2_bench_${N}gpu.log | cut -c1-4500
