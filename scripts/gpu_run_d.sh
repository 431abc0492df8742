#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== full-size tests"; timeout 1200 python -X faulthandler -m pytest tests/test_gpu_full_size.py -x -v 2>&1 | head -90 | tee gpurun_out/d_pytest_full.txt
echo "== other gpu tests"; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 | tee gpurun_out/d_pytest_parity.txt
echo "== numa probe"; timeout 300 python scripts/numa_probe.py 2>&1 | tee gpurun_out/d_numa.txt
echo "== bench parity"; RN_TIMING=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/d_bench_parity.json 2> gpurun_out/d_bench_parity.err; cat gpurun_out/d_bench_parity.json; grep -i "numa" gpurun_out/d_bench_parity.err | head
