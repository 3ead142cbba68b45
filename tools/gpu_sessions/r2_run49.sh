#!/bin/bash
# Round 2, GPU session 49 (1 GPU): last check of the committed state: full GPU suite, smoke(), bench lines c1 / c4 / c3
# (config 2 and the launch lists: session 45).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_49_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2_49_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_49_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2_49_smoke.log
for c in c1 c4 c3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_49_bench_$c.json 2> gpurun_out/r2_49_bench_$c.err
  echo "bench $c exit $?"; cut -c1-200 gpurun_out/r2_49_bench_$c.json
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_49_bench_reference.json 2> gpurun_out/r2_49_bench_reference.err
echo "reference arm exit $?"; cut -c1-300 gpurun_out/r2_49_bench_reference.json
