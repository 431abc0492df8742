#!/bin/bash
# cfg 3 (logreg 50 x 100k, 2048 chains) and cfg 2s (linreg 5 cov) with the chain-batched DMMA path on / off
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_mma.py -q -m gpu -x 2>&1 | tail -15
for mma in 1 0; do
  echo "== RN_MMA=$mma"
  RN_MMA=$mma timeout 900 python scripts/bench_configs.py cfg3 cfg2s --no-cpu --math=parity 2>&1 | cut -c1-700
done
