#!/bin/bash
# final validation session of round 1: every GPU test, smoke, bench lines, streamed configs, sanitizers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== gpu tests"; timeout 2400 python -X faulthandler -m pytest tests -m gpu -x -q --durations=5 2>&1 | grep -v "site-packages" | tail -40 | tee gpurun_out/m_pytest.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench parity"; timeout 600 python bench.py > gpurun_out/m_bench_parity.json 2> gpurun_out/m_bench_parity.err; cat gpurun_out/m_bench_parity.json; tail -3 gpurun_out/m_bench_parity.err
echo "== bench fast"; timeout 600 python bench.py --math fast --no-cpu-baseline > gpurun_out/m_bench_fast.json 2> gpurun_out/m_bench_fast.err; cut -c1-600 gpurun_out/m_bench_fast.json
echo "== configs"; timeout 600 python scripts/bench_configs.py cfg2 cfg2s cfg3 cfg4 cfg5 --no-cpu 2>&1 | cut -c1-330 | tee gpurun_out/m_configs.jsonl
echo "== racecheck (wpc)"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_probe.py wpc > gpurun_out/m_san_racecheck_wpc.txt 2>&1; echo "exit $?"; grep -E "RACECHECK SUMMARY|ok|Race reported" gpurun_out/m_san_racecheck_wpc.txt | head -12
