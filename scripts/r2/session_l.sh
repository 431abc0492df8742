#!/bin/bash
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== fast math tests"; timeout 900 python -m pytest tests/test_gpu_fast_math.py -q -m gpu 2>&1 | tail -4
for v in "4 4" "3 4" "2 4" "4 2" "6 2" "2 8"; do
  set -- $v
  echo "== cfg5 RN_WPC_WARPS=$1 RN_WPC_K=$2"; RN_WPC_WARPS=$1 RN_WPC_K=$2 timeout 900 python scripts/bench_configs.py cfg5 --no-cpu --math=parity 2>&1 | cut -c1-260
done
echo "== function / optimize benches"; timeout 600 python scripts/bench_function.py 2>&1 | tail -1 | cut -c1-600; timeout 600 python scripts/bench_optimize.py 2>&1 | tail -1 | cut -c1-600
