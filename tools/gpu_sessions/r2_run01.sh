#!/bin/bash
# Round 2, GPU session 1 (1 GPU): sanity of the round-1 state on this round's pod + the measurements round 1 left open:
# config 4 weak-scaling line at N=1, the chunked MT19937 generation test, one full-size (155 GB) ncu capture of the two
# X-stream kernels (why is the full problem 8-11 % slower than a 19 GB shard?), racecheck over the TMA-ring kernels.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_01_gpu.txt 2>&1
nproc >> gpurun_out/r2_01_gpu.txt; free -g >> gpurun_out/r2_01_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_01_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2_01_tests.log
B2_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k chunked_generation > gpurun_out/r2_01_chunked.log 2>&1
echo "chunked exit $?"; tail -3 gpurun_out/r2_01_chunked.log
timeout 600 python bench.py --config c4w --steps 10 --warmup 3 > gpurun_out/r2_01_bench_c4w_n1.json 2> gpurun_out/r2_01_bench_c4w_n1.err
echo "c4w exit $?"; cat gpurun_out/r2_01_bench_c4w_n1.json | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xstream_(y|z)" -s 2 -c 2 -o gpurun_out/r2_01_prof_xstream_full \
   python bench.py --config c2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_01_ncu_full.log 2>&1
echo "ncu full-size exit $?"; tail -3 gpurun_out/r2_01_ncu_full.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "xstream and not large" > gpurun_out/r2_01_racecheck_xstream.log 2>&1
echo "racecheck xstream exit $?"; tail -5 gpurun_out/r2_01_racecheck_xstream.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "rowpass" > gpurun_out/r2_01_racecheck_rowpass.log 2>&1
echo "racecheck rowpass exit $?"; tail -5 gpurun_out/r2_01_racecheck_rowpass.log
