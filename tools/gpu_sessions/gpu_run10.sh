#!/bin/bash
# Re-entry check of HEAD: full -m gpu suite, smoke, c2 bench + reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -25 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 300 python tools/phase_times.py > gpurun_out/phase_c2.log 2>&1; tail -30 gpurun_out/phase_c2.log
