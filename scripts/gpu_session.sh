#!/bin/bash
# One parameterised GPU session runner (replaces round 1's gpu_run_a..o.sh): each argument names a step; output of
# every step lands in gpurun_out/<tag>_<step>.log.  Usage (on the box, through gpurun):
#   bash scripts/gpu_session.sh <tag> step1 step2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RN_KERNEL_CACHE=$PWD/build/kcache
TAG=$1; shift
for step in "$@"; do
  log=gpurun_out/${TAG}_${step}.log
  echo "== $step"
  case $step in
    zz)        timeout 1200 python -m pytest tests/test_zz_gpu_function.py tests/test_zz_gpu_optimizer.py tests/test_zz_gpu_wpc_dense.py -q -m gpu -rA > $log 2>&1; tail -25 $log ;;
    diagfn)    timeout 600 python scripts/r2/diag_function.py > $log 2>&1; tail -40 $log ;;
    tests)     timeout 2400 python -m pytest tests -q -m gpu -rA > $log 2>&1; tail -15 $log ;;
    bench)     timeout 900 python bench.py > $log 2>&1; tail -2 $log | cut -c1-1500 ;;
    benchref)  timeout 900 python bench.py --impl reference > $log 2>&1; tail -1 $log | cut -c1-800 ;;
    configs)   timeout 1500 python scripts/bench_configs.py > $log 2>&1; tail -12 $log | cut -c1-600 ;;
    dmma)      timeout 300 python scripts/probe_dmma.py > $log 2>&1; tail -3 $log | cut -c1-1200 ;;
    launches)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > $log 2>&1; tail -2 $log | cut -c1-300 ;;
    ncufunnel) timeout 900 ncu --set full --clock-control none --import-source on -k regex:rn_k_iter -c 1 -o gpurun_out/${TAG}_ncu_funnel python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $log 2>&1; tail -2 $log | cut -c1-300 ;;
    sweep)     timeout 1200 python scripts/r2/sweep_iter.py > $log 2>&1; cat $log | cut -c1-300 ;;
    diagdense) timeout 900 python scripts/r2/diag_wpc_dense.py > $log 2>&1; tail -120 $log ;;
    *)         if [ -f "$step" ]; then timeout 1800 bash "$step" > gpurun_out/${TAG}_$(basename $step).log 2>&1; tail -30 gpurun_out/${TAG}_$(basename $step).log; else echo "unknown step $step"; fi ;;
  esac
done
