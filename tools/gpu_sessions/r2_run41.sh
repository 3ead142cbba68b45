#!/bin/bash
# Round 2, GPU session 41 (1 GPU): unimodal variant 14 as the default: unimodal tests (incl. the multi-round and
# few-long-groups cases), config-3 bench lines fp64 / fp32.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "unimodal" > gpurun_out/r2_41_tests_unimodal.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_41_tests_unimodal.log
for c in c3 c3f32; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_41_bench_$c.json 2> gpurun_out/r2_41_bench_$c.err
  echo "bench $c exit $?"; cut -c1-200 gpurun_out/r2_41_bench_$c.json
done
