#!/bin/bash
# Round 2, GPU session 45 (1 GPU): state of the round: full GPU suite, smoke(), default bench (config 2, full size, with
# e2e and the CPU arm), ncu launch lists of config 2 / config 3 at 2 048 slices (shares of the step).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_45_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2_45_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_45_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2_45_smoke.log
timeout 900 python bench.py > gpurun_out/r2_45_bench_c2.json 2> gpurun_out/r2_45_bench_c2.err
echo "bench c2 exit $?"; cut -c1-200 gpurun_out/r2_45_bench_c2.json
for c in c2 c3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_45_launches_${c}s.csv \
     python bench.py --config $c --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_45_launches_${c}s.log 2>&1
  echo "launch list $c exit $?"
done
