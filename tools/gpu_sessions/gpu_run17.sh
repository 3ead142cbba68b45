#!/bin/bash
# compute-sanitizer memcheck over the kernels changed late in the session (gap DMMA kernel, narrow-K Z contraction,
# sharded/random penalty combinations)
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "pf2_gap or (xstream_z and fma and f32) or unimodal_golden" > gpurun_out/memcheck_late_kernels.log 2>&1
echo "memcheck late kernels exit $?"; tail -5 gpurun_out/memcheck_late_kernels.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_aoadmm.py -m gpu -q -x -p no:cacheprovider -k "random_penalty or svd_inits or float32_inputs" > gpurun_out/memcheck_late_e2e.log 2>&1
echo "memcheck late e2e exit $?"; tail -5 gpurun_out/memcheck_late_e2e.log
