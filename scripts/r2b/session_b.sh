#!/bin/bash
# round 2b, session b: branch-free row functions (RN_ROW_LIBM) A/B on the streamed configurations, the warp-per-chain parity tests
# with them on, a fresh ncu capture of cfg 5, two more variants of the headline kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" timeout 600 python scripts/bench_configs.py $cfg --no-cpu --math=parity 2>&1 | grep '^{' | cut -c1-330; }
run cfg3 RN_ROW_LIBM=0
run cfg3 RN_ROW_LIBM=1
run cfg3 RN_ROW_LIBM=1 RN_MMA_ELEMS=2
run cfg5 RN_ROW_LIBM=0
run cfg5 RN_ROW_LIBM=1
run cfg5 RN_ROW_LIBM=1 RN_INTERLEAVE=2
run cfg2s RN_ROW_LIBM=0 RN_INLINE=0
run cfg2s RN_ROW_LIBM=1 RN_INLINE=0
echo "== row function probe + warp-per-chain parity tests with RN_ROW_LIBM=1"
RN_ROW_LIBM=1 timeout 900 python -m pytest tests/test_gpu_divsqrt.py tests/test_gpu_mma.py tests/test_gpu_full_size.py tests/test_gpu_fast_math.py tests/test_zz_gpu_wpc_dense.py "tests/test_gpu_parity.py" -q -m gpu -k "row_functions or mma or cfg3 or cfg5 or wpc or streamed or fast or dense or poisson or logistic" 2>&1 | tail -8
echo "== ncu cfg5 (RN_ROW_LIBM=1)"
RN_ROW_LIBM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2b_b_ncu_cfg5 python scripts/bench_configs.py cfg5 --no-cpu --math=parity > gpurun_out/r2b_b_ncu_cfg5.log 2>&1; tail -1 gpurun_out/r2b_b_ncu_cfg5.log | cut -c1-200
echo "== funnel variants"
SWEEP_CAPS="128" SWEEP_DEFS="|-DRN_X_NORMALS=0|-DRN_X_NORMALS=1" timeout 600 python scripts/r2/sweep_iter.py 2>/dev/null | cut -c1-200
