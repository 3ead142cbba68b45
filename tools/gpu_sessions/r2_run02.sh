#!/bin/bash
# Round 2, GPU session 2 (1 GPU): the new parity tests — BASELINE-width trajectories, 2/4/8-rank runs with gloo on one
# GPU, the non-SPD / zero-coordinate-matrix / empty-shard paths.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_baseline_widths.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_02_widths.log 2>&1
echo "widths exit $?"; grep "width parity\|passed\|failed\|Error\|assert" gpurun_out/r2_02_widths.log | tail -20
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_02_multi.log 2>&1
echo "multi exit $?"; grep "multi-rank parity\|passed\|failed\|Error" gpurun_out/r2_02_multi.log | tail -20
timeout 600 python -m pytest tests/test_gpu_aoadmm.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "indefinite or zero_coordinate or not_positive_definite or factor_batch" > gpurun_out/r2_02_new.log 2>&1
echo "new tests exit $?"; tail -15 gpurun_out/r2_02_new.log
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_baseline_widths.py --deselect tests/test_gpu_multi.py > gpurun_out/r2_02_tests.log 2>&1
echo "all other tests exit $?"; tail -3 gpurun_out/r2_02_tests.log
