#!/bin/bash
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 900 python -m pytest tests/test_gpu_aoadmm.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -5 gpurun_out/tests.log
timeout 900 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --config c2 --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "ref exit $?"; cat gpurun_out/bench_c2_ref.json
cat gpurun_out/host.txt
