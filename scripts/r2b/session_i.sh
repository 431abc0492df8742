#!/bin/bash
# round 2b, session i (last GPU seconds of the round): the row functions' default now depends on the register accumulators of the
# row body -- the rows-across-lanes form of cfg 3 must be back at its round-2 rate, cfg 5 unchanged; streamed parity tests
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" timeout 100 python scripts/bench_configs.py $cfg --no-cpu --math=parity 2>&1 | grep '^{' | cut -c1-300; }
run cfg3 RN_MMA=0
run cfg5 RN_DUMMY=1
timeout 100 python -m pytest tests/test_gpu_mma.py tests/test_gpu_parity.py -q -m gpu -x -k "mma or wpc_logistic or wpc_poisson or streamed_targets" 2>&1 | tail -3
