#!/bin/bash
# compute-sanitizer over the kernels added in this session (memcheck everywhere, racecheck on the warp-level kernels)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "next_penalties or tv_kernel or polar_warp_vs_cta or single_array or prox_options or polar_warm" > gpurun_out/memcheck_kernels.log 2>&1
echo "memcheck kernels exit $?"; tail -6 gpurun_out/memcheck_kernels.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_aoadmm.py -m gpu -q -x -p no:cacheprovider -k "trajectory and (gl2 or simplex or tv_ or pf2_ or c2_nn or c3_)" > gpurun_out/memcheck_e2e.log 2>&1
echo "memcheck e2e exit $?"; tail -6 gpurun_out/memcheck_e2e.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "polar_warp_vs_cta and (f64-3 or f64-8 or f64-20 or f64-32)" > gpurun_out/racecheck_polar.log 2>&1
echo "racecheck polar exit $?"; tail -12 gpurun_out/racecheck_polar.log
