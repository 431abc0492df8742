#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== cfg3 default (W=8, 2 stages, 255 regs)"; timeout 600 python scripts/bench_configs.py cfg3 --no-cpu 2>&1 | tail -1 | cut -c1-420
echo "== cfg3 W=10, 2 stages, 200 regs"; RN_WPC_WARPS=12 timeout 600 python scripts/bench_configs.py cfg3 --no-cpu 2>&1 | tail -1 | cut -c1-420
echo "== cfg3 W=12, 1 stage, 168 regs"; RN_TMA=1 RN_WPC_WARPS=12 timeout 600 python scripts/bench_configs.py cfg3 --no-cpu 2>&1 | tail -1 | cut -c1-420
echo "== cfg3 W=16, 1 stage, 128 regs"; RN_TMA=1 RN_WPC_WARPS=16 timeout 600 python scripts/bench_configs.py cfg3 --no-cpu 2>&1 | tail -1 | cut -c1-420
echo "== cfg5 auto (K=4)"; timeout 600 python scripts/bench_configs.py cfg5 --no-cpu 2>&1 | tail -1 | cut -c1-420
