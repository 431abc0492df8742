#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== gpu tests"; timeout 1800 python -X faulthandler -m pytest tests -m gpu -x -q --durations=6 2>&1 | grep -v "site-packages" | head -60 | tee gpurun_out/e_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench parity"; timeout 600 python bench.py > gpurun_out/e_bench_parity.json 2> gpurun_out/e_bench_parity.err; cat gpurun_out/e_bench_parity.json; tail -3 gpurun_out/e_bench_parity.err
echo "== bench fast"; timeout 600 python bench.py --math fast --no-cpu-baseline > gpurun_out/e_bench_fast.json 2> gpurun_out/e_bench_fast.err; cat gpurun_out/e_bench_fast.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/e_bench_reference.json | cut -c1-400
