#!/bin/bash
# after: ld.shared tile reads, register error flag, native D2I, adj*(1/x) reverse-sweep quotients
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
run() { cfg=$1; shift; echo "== $cfg $*"; env "$@" timeout 900 python scripts/bench_configs.py $cfg --no-cpu --math=parity 2>&1 | tail -1 | cut -c1-330; }
run cfg5 RN_DUMMY=1
run cfg3 RN_DUMMY=1
run cfg3 RN_EXACT_ROW_DIV=1
run cfg3 RN_NO_REROLL=1
run cfg3 RN_MMA=0
echo "== all gpu tests"; timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --no-configs 2>&1 | tail -1 | cut -c1-1800
echo "== ncu cfg3"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/r2n_ncu_cfg3 python scripts/bench_configs.py cfg3 --no-cpu --math=parity > gpurun_out/r2n_ncu_cfg3.log 2>&1; tail -1 gpurun_out/r2n_ncu_cfg3.log | cut -c1-200
