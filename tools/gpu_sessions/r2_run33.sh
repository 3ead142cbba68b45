#!/bin/bash
# Round 2, GPU session 33 (1 GPU): full GPU suite + smoke() + the driver-style default bench on the state after the
# 12-warp Y / 16-warp Z X-stream kernels, the small-kernel work and the CUDA-graph launch policy.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_33_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_33_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_33_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/r2_33_smoke.log
timeout 900 python bench.py > gpurun_out/r2_33_bench_c2.json 2> gpurun_out/r2_33_bench_c2.err
echo "bench c2 exit $?"; cut -c1-200 gpurun_out/r2_33_bench_c2.json
timeout 600 python bench.py --config c3 --steps 20 --warmup 5 > gpurun_out/r2_33_bench_c3.json 2> gpurun_out/r2_33_bench_c3.err
echo "bench c3 exit $?"; cut -c1-200 gpurun_out/r2_33_bench_c3.json
