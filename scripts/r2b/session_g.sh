#!/bin/bash
# round 2b, session g: validation after the last two changes of the headline kernel (state kept in registers, compile-time CTA size)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== bench"; timeout 600 python bench.py > gpurun_out/r2_final2_bench.log 2>&1; tail -1 gpurun_out/r2_final2_bench.log | cut -c1-1800
echo "== all gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -rA > gpurun_out/r2_final2_tests.log 2>&1; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2_final2_tests.log | tail -12 | cut -c1-200
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== ncu funnel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2_final2_ncu_funnel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_final2_ncu_funnel.log 2>&1; tail -1 gpurun_out/r2_final2_ncu_funnel.log | cut -c1-160
timeout 300 python scripts/ncu_summary.py gpurun_out/r2_final2_ncu_funnel.ncu-rep r2_ncu_funnel_parity_v4 parity 2>&1 | tail -1 | cut -c1-500
cp profiles/r2_ncu_funnel_parity_v4*.csv profiles/ncu_funnel_parity.json gpurun_out/ 2>/dev/null
echo "== launches"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final2_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/r2_final2_launches.log 2>&1; tail -1 gpurun_out/r2_final2_launches.log | cut -c1-160
echo "== configs"; timeout 600 python scripts/bench_configs.py cfg2 cfg4 --math=parity --no-cpu 2>&1 | grep '^{' | cut -c1-330
