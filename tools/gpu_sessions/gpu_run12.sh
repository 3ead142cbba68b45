#!/bin/bash
# register polar kernel + T-only companions: full -m gpu suite, polar micro-bench, per-op phase times at c2 / c3, c2 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -12 gpurun_out/tests.log
timeout 300 python tools/bench_polar.py 2>&1 | grep reg
timeout 300 python tools/phase_times.py > gpurun_out/phase_c2.log 2>&1; head -8 gpurun_out/phase_c2.log
timeout 300 python tools/phase_times.py c3 > gpurun_out/phase_c3.log 2>&1; head -8 gpurun_out/phase_c3.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; cut -c1-300 gpurun_out/bench_c2.json
