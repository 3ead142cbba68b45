#!/bin/bash
# Round 2, GPU session 42 (1 GPU): unimodal kernel with compact prefix errors for the peak search (variant 18: top block updated in place):
# bit-exactness tests under every variant, A/B at config-3 size.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "unimodal" > gpurun_out/r2_42_tests_unimodal.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_42_tests_unimodal.log
timeout 600 python tools/ab_unimodal.py > gpurun_out/r2_42_ab_unimodal.log 2>&1
echo "ab exit $?"; cat gpurun_out/r2_42_ab_unimodal.log | cut -c1-220
