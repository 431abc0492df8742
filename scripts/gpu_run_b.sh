#!/bin/bash
# Round-1 GPU session B: new streamed-kernel emission (+TMA tiles), pipelined rn_sample, diagnostics; A/B switches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/b_pytest_gpu.txt
echo "== bench parity"; timeout 600 python bench.py > gpurun_out/b_bench_parity.json 2> gpurun_out/b_bench_parity.err; cat gpurun_out/b_bench_parity.json; tail -3 gpurun_out/b_bench_parity.err
echo "== bench parity, fdlibm inlined"; RN_LIBM_INLINE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_bench_parity_inline.json 2>&1; cut -c1-330 gpurun_out/b_bench_parity_inline.json
echo "== bench fast"; timeout 600 python bench.py --math fast --no-cpu-baseline > gpurun_out/b_bench_fast.json 2> gpurun_out/b_bench_fast.err; cat gpurun_out/b_bench_fast.json
echo "== e2e probe (pageable)"; RN_TIMING=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -16 | tee gpurun_out/b_e2e_probe.txt
echo "== configs default"; timeout 900 python scripts/bench_configs.py cfg2s cfg3 cfg5 > gpurun_out/b_configs.jsonl 2> gpurun_out/b_configs.err; cat gpurun_out/b_configs.jsonl; tail -5 gpurun_out/b_configs.err
echo "== configs RN_TMA=0"; RN_TMA=0 timeout 900 python scripts/bench_configs.py cfg2s cfg3 --no-cpu > gpurun_out/b_configs_tma0.jsonl 2> gpurun_out/b_configs_tma0.err; cat gpurun_out/b_configs_tma0.jsonl; tail -5 gpurun_out/b_configs_tma0.err
echo "== configs RN_TMA=1 cfg3"; RN_TMA=1 timeout 900 python scripts/bench_configs.py cfg3 --no-cpu > gpurun_out/b_configs_tma1.jsonl 2> gpurun_out/b_configs_tma1.err; cat gpurun_out/b_configs_tma1.jsonl; tail -5 gpurun_out/b_configs_tma1.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/b_launches_bench_funnel.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_launches.log 2>&1
echo "== ncu full cfg3 fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 1 -c 1 -o gpurun_out/b_prof_cfg3_fast \
  python scripts/bench_configs.py cfg3 --math=fast --no-cpu > gpurun_out/b_ncu_cfg3_fast.log 2>&1
ls -la gpurun_out | grep " b_"
