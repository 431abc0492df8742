#!/bin/bash
# round 2b, session a: check-free division / sqrt + speculative fdlibm paths + constant-bank coefficients on the device:
# bit-for-bit probes, the A/B sweep of the headline kernel (sample hashes must not move), then every GPU test
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== divsqrt probe tests"; timeout 600 python -m pytest tests/test_gpu_divsqrt.py -q -m gpu -x 2>&1 | tail -15
echo "== sweep"
SWEEP_CAPS="128,96" SWEEP_DEFS="-DRN_X_NOCHECK_DIV=0 -DRN_X_SPEC=0 -DRN_X_KCONST=0 -DRN_X_POLAR_FMA=0|-DRN_X_SPEC=0 -DRN_X_KCONST=0 -DRN_X_POLAR_FMA=0|-DRN_X_KCONST=0 -DRN_X_POLAR_FMA=0|-DRN_X_POLAR_FMA=0||-DRN_X_NORMALS=4|-DRN_X_NORMALS=3" \
  timeout 900 python scripts/r2/sweep_iter.py > gpurun_out/r2b_a_sweep.jsonl 2> gpurun_out/r2b_a_sweep.err; cut -c1-260 gpurun_out/r2b_a_sweep.jsonl; tail -3 gpurun_out/r2b_a_sweep.err
echo "== all gpu tests"; timeout 1000 python -m pytest tests -q -m gpu -rA > gpurun_out/r2b_a_tests.log 2>&1; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r2b_a_tests.log | tail -20 | cut -c1-220
