#!/bin/bash
# Round 2, GPU session 52 (1 GPU): compute-sanitizer memcheck over the unimodal kernels of the last session (compact
# error copy, deferred fill kernel, multi-round fallback) and the trajectories that use them.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "unimodal" > gpurun_out/r2_52_memcheck_unimodal.log 2>&1
echo "memcheck kernels exit $?"; tail -4 gpurun_out/r2_52_memcheck_unimodal.log
