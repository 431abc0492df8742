#!/bin/bash
# round 2b, session c: opt-in variants of the headline kernel (two-attempt polar pass, merged-fallback density), cfg 5 with the
# new defaults (row functions on, lookup index reused by the scatter, two observations in flight) against each switch turned back
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== funnel variants"
SWEEP_CAPS="128" SWEEP_DEFS="|-DRN_X_POLAR2=1|ENV:RN_MERGED_FALLBACK=1|-DRN_X_POLAR2=1 ENV:RN_MERGED_FALLBACK=1|-DRN_X_POLAR2=1 -DRN_X_NORMALS=4" timeout 600 python scripts/r2/sweep_iter.py 2>/dev/null | cut -c1-220 | tee gpurun_out/r2b_c_sweep.jsonl
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" timeout 600 python scripts/bench_configs.py $cfg --no-cpu --math=parity 2>&1 | grep '^{' | cut -c1-330; }
run cfg5 RN_DUMMY=1
run cfg5 RN_SCATTER_REUSE_INDEX=0
run cfg5 RN_INTERLEAVE=4
run cfg5 RN_INTERLEAVE=1
run cfg5 RN_ROW_LIBM=0
run cfg3 RN_DUMMY=1
run cfg4 RN_DUMMY=1
run cfg2 RN_DUMMY=1
echo "== parity tests of the streamed shapes with the new defaults"
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_fast_math.py tests/test_gpu_parity.py -q -m gpu -k "cfg5 or wpc or streamed or poisson or scatter" 2>&1 | tail -5
