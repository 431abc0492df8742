#!/bin/bash
# Session O (first session of round 2): GPU confirmation + measurement of the rows added CPU-only at the end of round 1
# (8f-2 rn_function_*, 8f-4 rn_optimize, dense mass on the warp-per-chain kernels), then the standing evidence (tests, bench, launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== pytest -m gpu (new rows first)"; timeout 900 python -m pytest tests/test_zz_gpu_function.py tests/test_zz_gpu_optimizer.py tests/test_zz_gpu_wpc_dense.py -q -m gpu 2>&1 | tail -15
echo "== bench_function parity"; timeout 300 python scripts/bench_function.py | tee gpurun_out/bench_function_parity.json | cut -c1-1500
echo "== bench_function fast"; timeout 300 python scripts/bench_function.py --fast | tee gpurun_out/bench_function_fast.json | cut -c1-600
echo "== bench_optimize parity"; timeout 300 python scripts/bench_optimize.py | tee gpurun_out/bench_optimize_parity.json | cut -c1-1500
echo "== ncu rn_k_eval"; timeout 600 ncu --set full --clock-control none --import-source on -k rn_k_eval -c 1 -o gpurun_out/ncu_rn_k_eval python scripts/bench_function.py --steps 1 --warmup 0 > gpurun_out/ncu_eval.log 2>&1; tail -3 gpurun_out/ncu_eval.log
echo "== ncu rn_k_lbfgs"; timeout 600 ncu --set full --clock-control none --import-source on -k rn_k_lbfgs -c 1 -o gpurun_out/ncu_rn_k_lbfgs python scripts/bench_optimize.py --steps 1 --warmup 0 --starts 37888 > gpurun_out/ncu_lbfgs.log 2>&1; tail -3 gpurun_out/ncu_lbfgs.log
echo "== full pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py | tee gpurun_out/bench_parity.json | cut -c1-400
echo "== compute-sanitizer memcheck / racecheck on the new kernels"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_function.py -q -m gpu -k "bit_identical and not 300000" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_optimizer.py -q -m gpu -k "warp_per_start" 2>&1 | tail -4
echo "== DMMA probe (fragment mapping + fp64 tensor-core rate at cfg 3's shape)"; timeout 300 python scripts/probe_dmma.py | tee gpurun_out/probe_dmma.json | cut -c1-800
