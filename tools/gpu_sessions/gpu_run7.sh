#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -25 gpurun_out/tests.log
for c in c2 c3; do timeout 300 python tools/phase_times.py $c 2>&1 | head -9; done
timeout 900 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -3 gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
