#!/bin/bash
# round 2b, final validation of the committed state: ncu capture of the headline kernel (and its summary, which bench.py quotes),
# every GPU test, smoke(), the bench line + the reference arm, the launch list of bench.py, the BASELINE configurations
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== ncu funnel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2_final_ncu_funnel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_final_ncu_funnel.log 2>&1; tail -1 gpurun_out/r2_final_ncu_funnel.log | cut -c1-200
timeout 300 python scripts/ncu_summary.py gpurun_out/r2_final_ncu_funnel.ncu-rep r2_ncu_funnel_parity_v3 parity 2>&1 | tail -2 | cut -c1-600
cp profiles/r2_ncu_funnel_parity_v3*.csv profiles/ncu_funnel_parity.json gpurun_out/ 2>/dev/null
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -rA --durations=10 > gpurun_out/r2_final_tests.log 2>&1; grep -E "FAILED|ERROR|passed|failed|^[0-9.]+s " gpurun_out/r2_final_tests.log | tail -16 | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > gpurun_out/r2_final_bench.log 2>&1; tail -1 gpurun_out/r2_final_bench.log | cut -c1-2500
echo "== bench --impl reference"; timeout 300 python bench.py --impl reference > gpurun_out/r2_final_bench_reference.log 2>&1; tail -1 gpurun_out/r2_final_bench_reference.log | cut -c1-700
echo "== launches"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2_final_launches.log 2>&1; tail -1 gpurun_out/r2_final_launches.log | cut -c1-160
echo "== configs"; timeout 700 python scripts/bench_configs.py --math=parity > gpurun_out/r2_final_configs.log 2>&1; grep '^{' gpurun_out/r2_final_configs.log | cut -c1-420
