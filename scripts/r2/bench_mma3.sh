#!/bin/bash
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
for nc in 16 8; do
  echo "== RN_MMA_CHAINS=$nc"; RN_MMA_CHAINS=$nc timeout 900 python scripts/bench_configs.py cfg3 --no-cpu --math=parity 2>&1 | cut -c1-330
done
echo "== tests"; timeout 1800 python -m pytest tests/test_gpu_mma.py tests/test_gpu_full_size.py -q -m gpu 2>&1 | tail -8
echo "== goldsets"; timeout 1800 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "goldsets or host_buffers" 2>&1 | tail -8
echo "== ncu cfg3"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2i_ncu_cfg3 python scripts/bench_configs.py cfg3 --no-cpu --math=parity > gpurun_out/r2i_ncu_cfg3.log 2>&1; tail -2 gpurun_out/r2i_ncu_cfg3.log | cut -c1-200
