#!/bin/bash
# round 2b, session f: state kept in registers across accepted iterations, compile-time block size (headline kernel A/B)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
SWEEP_CAPS="128" SWEEP_DEFS="|-DRN_X_KEEP_STATE=1|-DRN_BLOCK_DIM=128|-DRN_BLOCK_DIM=128 -DRN_X_KEEP_STATE=1" timeout 600 python scripts/r2/sweep_iter.py 2>/dev/null | cut -c1-220 | tee gpurun_out/r2b_f_sweep.jsonl
SWEEP_CAPS="112,104" SWEEP_DEFS="-DRN_BLOCK_DIM=128 -DRN_X_KEEP_STATE=1" timeout 600 python scripts/r2/sweep_iter.py 2>/dev/null | cut -c1-220 | tee -a gpurun_out/r2b_f_sweep.jsonl
