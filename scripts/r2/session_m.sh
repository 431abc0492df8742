#!/bin/bash
# cfg 5 after the re-rolled invariant sections, the dropped mass slice and the second data-tile stage
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
run() { echo "== cfg5 $*"; env "$@" timeout 900 python scripts/bench_configs.py cfg5 --no-cpu --math=parity 2>&1 | tail -1 | cut -c1-330; }
run RN_DUMMY=1
run RN_NO_REROLL=1
run RN_TMA=1
run RN_INTERLEAVE=8
run RN_WPC_WARPS=3
echo "== cfg5 fast"; timeout 900 python scripts/bench_configs.py cfg5 --no-cpu --math=fast 2>&1 | tail -1 | cut -c1-330
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_fast_math.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5
echo "== cfg3 / cfg4 / cfg2s"; timeout 900 python scripts/bench_configs.py cfg3 cfg4 cfg2s --no-cpu --math=parity 2>&1 | grep '^{' | cut -c1-330
echo "== ncu cfg5"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2m_ncu_cfg5 python scripts/bench_configs.py cfg5 --no-cpu --math=parity > gpurun_out/r2m_ncu_cfg5.log 2>&1; tail -1 gpurun_out/r2m_ncu_cfg5.log | cut -c1-200
