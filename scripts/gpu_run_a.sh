#!/bin/bash
# Round-1 GPU session A: parity tests, headline bench (both math modes), PCIe/drain probes, other configs, ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench parity"; timeout 600 python bench.py > gpurun_out/bench_parity.json 2> gpurun_out/bench_parity.err; tail -c 3000 gpurun_out/bench_parity.json
echo "== bench fast"; timeout 600 python bench.py --math fast --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 1500 gpurun_out/bench_fast.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2>&1
echo "== pcie probe"; timeout 300 python scripts/pcie_probe.py > gpurun_out/pcie_probe.txt 2>&1; cat gpurun_out/pcie_probe.txt
echo "== e2e probe"; RN_TIMING=1 timeout 300 python scripts/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1; tail -40 gpurun_out/e2e_probe.txt
echo "== configs"; timeout 1500 python scripts/bench_configs.py cfg2 cfg2s cfg4 cfg3 cfg5 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err; cat gpurun_out/bench_configs.jsonl; tail -5 gpurun_out/bench_configs.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench_funnel.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "== ncu full rn_k_iter"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 3 -c 1 -o gpurun_out/prof_funnel_parity \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 3 -c 1 -o gpurun_out/prof_funnel_fast \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --math fast > gpurun_out/ncu_full_fast.log 2>&1
ls -la gpurun_out
echo "== ncu full cfg3 / cfg5 (fast)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 1 -c 1 -o gpurun_out/prof_cfg3_fast \
  python scripts/bench_configs.py cfg3 --math=fast --no-cpu > gpurun_out/ncu_cfg3_fast.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -s 1 -c 1 -o gpurun_out/prof_cfg5_fast \
  python scripts/bench_configs.py cfg5 --math=fast --no-cpu > gpurun_out/ncu_cfg5_fast.log 2>&1
ls -la gpurun_out
