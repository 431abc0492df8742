#!/bin/bash
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== inline tests"; timeout 900 python -m pytest tests/test_gpu_inline.py tests/test_gpu_mma.py -q -m gpu 2>&1 | tail -6
echo "== cfg3 / cfg2s / cfg2si"; timeout 900 python scripts/bench_configs.py cfg3 cfg2si cfg2s --no-cpu --math=parity 2>&1 | cut -c1-330
echo "== funnel sweep"; SWEEP_DEFS="" SWEEP_CAPS=128 timeout 600 python scripts/r2/sweep_iter.py 2>&1 | cut -c1-200
echo "== bench"; timeout 900 python bench.py > gpurun_out/r2j_bench.log 2>&1; tail -1 gpurun_out/r2j_bench.log | cut -c1-3000
