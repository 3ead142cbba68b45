#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host.txt; nproc >> gpurun_out/host.txt; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv >> gpurun_out/host.txt
timeout 300 python tools/debug_z.py > gpurun_out/debug_z.log 2>&1; tail -20 gpurun_out/debug_z.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/kernels.log; tail -15 gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_gpu_aoadmm.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/aoadmm.log 2>&1
echo "aoadmm exit $?" >> gpurun_out/aoadmm.log; tail -15 gpurun_out/aoadmm.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log; grep -E "xstream|outer|step_|refresh|diag|exit|Error|error" gpurun_out/probe.log | tail -60
timeout 600 python bench.py --config c1 --steps 10 --warmup 3 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 exit $?"; tail -3 gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json
timeout 900 python bench.py --config c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2s.csv python bench.py --config c2 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c2s.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3s.csv python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3s.log 2>&1; echo "ncu exit $?"
