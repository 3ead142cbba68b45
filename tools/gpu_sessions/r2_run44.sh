#!/bin/bash
# Round 2, GPU session 44 (re-run of the tests of session 43 after the graph-mode fix) (1 GPU): Jacobi warm start across B-updates (cold every 8th): PARAFAC2 trajectory parity,
# A/B of the polar step at 4 096 slices of config 2 (B2_POLAR_COLD_EVERY=1 is the previous behaviour).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_aoadmm.py tests/test_gpu_baseline_widths.py tests/test_gpu_penalty_contract.py tests/test_gpu_reference_kats.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_44_tests_aoadmm.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_44_tests_aoadmm.log
