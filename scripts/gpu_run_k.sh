#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== diagnostics tests"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "diagnostics or ragged or pinned" 2>&1 | grep -v site-packages | tail -15
echo "== memcheck (tpc + diag)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py tpc > gpurun_out/k_san_memcheck_tpc.txt 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|ok|Error" gpurun_out/k_san_memcheck_tpc.txt | head
echo "== bench parity"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/k_bench_parity.json 2> gpurun_out/k_bench_parity.err; python -c "
import json; d=json.load(open('gpurun_out/k_bench_parity.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_call'], d['e2e']['pageable_ms_per_call'], d['e2e']['diagnostics_only'])"; tail -3 gpurun_out/k_bench_parity.err
