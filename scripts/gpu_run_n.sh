#!/bin/bash
cd "$(dirname "$0")/.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== row exp = CUDA libm (default)"; timeout 600 python scripts/bench_configs.py cfg3 cfg5 --no-cpu --math=parity 2>&1 | cut -c1-300
echo "== row exp = fdlibm"; RN_ROW_EXP_FDLIBM=1 timeout 600 python scripts/bench_configs.py cfg3 cfg5 --no-cpu --math=parity 2>&1 | cut -c1-300
