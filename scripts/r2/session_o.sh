#!/bin/bash
# 2 GPUs: the optimizer test that failed in session n, then the multi-GPU checks
cd "$(dirname "$0")/../.."
export RN_KERNEL_CACHE=$PWD/build/kcache
echo "== optimizer tests"; timeout 900 python -m pytest tests/test_zz_gpu_optimizer.py tests/test_zz_gpu_function.py tests/test_gpu_inline.py tests/test_gpu_mma.py -q -m gpu 2>&1 | tail -4
bash scripts/r2/session_multi.sh 2
