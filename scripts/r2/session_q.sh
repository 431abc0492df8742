#!/bin/bash
# re-entry validation at HEAD: every GPU test (no -x), the bench line, its ncu launch list, the BASELINE configs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== all gpu tests"; timeout 1000 python -m pytest tests -q -m gpu -rA --durations=15 > gpurun_out/r2q_tests.log 2>&1; tail -30 gpurun_out/r2q_tests.log | cut -c1-200
echo "== bench"; timeout 400 python bench.py > gpurun_out/r2q_bench.log 2>&1; tail -1 gpurun_out/r2q_bench.log | cut -c1-3000
echo "== launches"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2q_launches.log 2>&1; tail -1 gpurun_out/r2q_launches.log | cut -c1-200
echo "== configs"; timeout 700 python scripts/bench_configs.py --math=parity > gpurun_out/r2q_configs.log 2>&1; grep '^{' gpurun_out/r2q_configs.log | cut -c1-400
