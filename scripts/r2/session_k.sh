#!/bin/bash
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== funnel sweep"; SWEEP_DEFS="|-DRN_X_NORMALS=3" SWEEP_CAPS=128 SWEEP_BLOCKS=128,96,64 timeout 600 python scripts/r2/sweep_iter.py 2>&1 | cut -c1-200
echo "== fast math tests"; timeout 900 python -m pytest tests/test_gpu_fast_math.py -q -m gpu 2>&1 | tail -8
echo "== ncu cfg5"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2k_ncu_cfg5 python scripts/bench_configs.py cfg5 --no-cpu --math=parity > gpurun_out/r2k_ncu_cfg5.log 2>&1; tail -2 gpurun_out/r2k_ncu_cfg5.log | cut -c1-300
