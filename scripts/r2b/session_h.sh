#!/bin/bash
# round 2b, session h: forward + reverse statements of a group of observations issued together (RN_ROW_FUSED_SWEEPS) -- A/B on the
# rows-across-lanes kernels, and their parity tests with the switch on
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" timeout 300 python scripts/bench_configs.py $cfg --no-cpu --math=parity 2>&1 | grep '^{' | cut -c1-300; }
run cfg5 RN_DUMMY=1
run cfg5 RN_ROW_FUSED_SWEEPS=1
run cfg5 RN_ROW_FUSED_SWEEPS=1 RN_INTERLEAVE=4
run cfg5 RN_ROW_FUSED_SWEEPS=1 RN_INTERLEAVE=1
run cfg5 RN_ROW_FUSED_SWEEPS=1 RN_INTERLEAVE=8
run cfg2s RN_DUMMY=1
run cfg2s RN_ROW_FUSED_SWEEPS=1
run cfg3 RN_MMA=0
run cfg3 RN_MMA=0 RN_ROW_FUSED_SWEEPS=1
echo "== parity tests with RN_ROW_FUSED_SWEEPS=1"
RN_ROW_FUSED_SWEEPS=1 timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py tests/test_gpu_fast_math.py -q -m gpu -k "cfg5 or wpc or streamed or poisson or scatter or laplace or goldsets" 2>&1 | tail -4
