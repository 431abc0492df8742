#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== adjoint-mode tests"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -k "adjoint or wpc or cfg3 or cfg5 or several" 2>&1 | grep -v site-packages | tail -8
echo "== configs"; timeout 600 python scripts/bench_configs.py cfg2s cfg3 cfg5 --no-cpu 2>&1 | cut -c1-330 | tee gpurun_out/l_configs.jsonl
echo "== funnel adjoint gradient (parity)"; timeout 300 python bench.py --grad adjoint --no-cpu-baseline --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e'%d['value'], 'ms %.3f'%d['ms_per_step'])"
