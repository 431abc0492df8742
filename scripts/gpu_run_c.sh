#!/bin/bash
# Round-1 GPU session C: full-size parity tests, interleaved row bodies, final bench lines + ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== pytest -m gpu" ; timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 | tee gpurun_out/c_pytest_gpu.txt
echo "== bench parity"; timeout 600 python bench.py > gpurun_out/c_bench_parity.json 2> gpurun_out/c_bench_parity.err; cat gpurun_out/c_bench_parity.json; tail -3 gpurun_out/c_bench_parity.err
echo "== bench fast"; timeout 600 python bench.py --math fast --no-cpu-baseline > gpurun_out/c_bench_fast.json 2> gpurun_out/c_bench_fast.err; cat gpurun_out/c_bench_fast.json
echo "== configs"; timeout 900 python scripts/bench_configs.py cfg2s cfg3 cfg5 --no-cpu > gpurun_out/c_configs.jsonl 2> gpurun_out/c_configs.err; cat gpurun_out/c_configs.jsonl; tail -5 gpurun_out/c_configs.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c_launches_bench_funnel.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu_launches.log 2>&1
echo "== ncu full funnel parity"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 3 -c 1 -o gpurun_out/c_prof_funnel_parity \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_ncu_full.log 2>&1
echo "== ncu full cfg3 fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 1 -c 1 -o gpurun_out/c_prof_cfg3_fast \
  python scripts/bench_configs.py cfg3 --math=fast --no-cpu > gpurun_out/c_ncu_cfg3_fast.log 2>&1
ls -la gpurun_out | grep " c_"
