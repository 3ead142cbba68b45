#!/bin/bash
# first GPU call: kernel parity tests (all, no -x), e2e parity, probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/kernels.log
tail -40 gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_gpu_aoadmm.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/aoadmm.log 2>&1
echo "aoadmm exit $?" >> gpurun_out/aoadmm.log
tail -40 gpurun_out/aoadmm.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
echo "probe exit $?" >> gpurun_out/probe.log
tail -60 gpurun_out/probe.log
