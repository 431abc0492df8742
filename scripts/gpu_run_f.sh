#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== K-warps tests"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "several_warps or wpc or pinned" 2>&1 | grep -v site-packages | tail -25 | tee gpurun_out/f_pytest.txt
echo "== cfg5 K=auto"; timeout 600 python scripts/bench_configs.py cfg5 --no-cpu 2>&1 | tail -2 | tee gpurun_out/f_cfg5_auto.jsonl
echo "== cfg5 K=4"; RN_WPC_K=4 timeout 600 python scripts/bench_configs.py cfg5 --no-cpu 2>&1 | tail -2 | tee gpurun_out/f_cfg5_k4.jsonl
echo "== cfg5 K=1"; RN_WPC_K=1 timeout 600 python scripts/bench_configs.py cfg5 --no-cpu 2>&1 | tail -2 | tee gpurun_out/f_cfg5_k1.jsonl
echo "== cfg3/cfg2s"; timeout 600 python scripts/bench_configs.py cfg2s cfg3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/f_cfg23.jsonl
echo "== bench parity (e2e stability)"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/f_bench_parity.json 2> gpurun_out/f_bench_parity.err; python -c "
import json; d=json.load(open('gpurun_out/f_bench_parity.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_call'], d['e2e']['pageable_ms_per_call'])"
echo "== regs sweep (funnel parity)"; for r in 112 96 80; do echo -n "maxreg=$r: "; RN_MAXRREGCOUNT=$r timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e'%d['value'], 'ms %.3f'%d['ms_per_step'])"; done
for b in 64 256; do echo -n "block=$b: "; RN_BLOCK=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e'%d['value'], 'ms %.3f'%d['ms_per_step'])"; done
