#!/bin/bash
# warp polar kernel + new penalties: full -m gpu suite, polar micro-bench, per-op phase times at c2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -40 gpurun_out/tests.log
timeout 300 python tools/bench_polar.py 2>&1 | tail -8
timeout 300 python tools/phase_times.py > gpurun_out/phase_c2.log 2>&1; head -8 gpurun_out/phase_c2.log
