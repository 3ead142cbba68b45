#!/bin/bash
# Round 2, GPU session 5 (1 GPU): one-array companion + fused gap terms, unimodal kernel variants (A/B), benches c2 / c3.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_aoadmm.py tests/test_gpu_baseline_widths.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_05_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_05_tests.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_reference_kats.py tests/test_gpu_penalty_contract.py -m gpu -q -x -p no:cacheprovider -k "not many_rank" > gpurun_out/r2_05_tests2.log 2>&1
echo "tests2 exit $?"; tail -3 gpurun_out/r2_05_tests2.log
timeout 600 python tools/ab_unimodal.py > gpurun_out/r2_05_ab_unimodal.log 2>&1
echo "ab unimodal exit $?"; grep -c variant gpurun_out/r2_05_ab_unimodal.log; tail -3 gpurun_out/r2_05_ab_unimodal.log
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_05_bench_c2.json 2> gpurun_out/r2_05_bench_c2.err
echo "bench c2 exit $?"; cut -c1-300 gpurun_out/r2_05_bench_c2.json; tail -3 gpurun_out/r2_05_bench_c2.err
timeout 900 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_05_bench_c3.json 2> gpurun_out/r2_05_bench_c3.err
echo "bench c3 exit $?"; cut -c1-300 gpurun_out/r2_05_bench_c3.json; tail -3 gpurun_out/r2_05_bench_c3.err
