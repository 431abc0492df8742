#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer $tool (wpc)"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_probe.py wpc > gpurun_out/i_san_${tool}_wpc.txt 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|error" gpurun_out/i_san_${tool}_wpc.txt | head -12
done
echo "== compute-sanitizer memcheck (tpc)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py tpc > gpurun_out/i_san_memcheck_tpc.txt 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|ok|Error|error" gpurun_out/i_san_memcheck_tpc.txt | head
echo "== new edge tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "degenerate or ragged" 2>&1 | tail -4
echo "== cfg2 cfg4 with the small-batch launch heuristic"; timeout 600 python scripts/bench_configs.py cfg2 cfg4 --no-cpu 2>&1 | cut -c1-330
