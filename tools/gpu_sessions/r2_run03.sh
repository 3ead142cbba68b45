#!/bin/bash
# Round 2, GPU session 3 (1 GPU): the steady-state row-pass kernel (bit-identity vs the general kernel, A/B timing,
# ncu), the 8-rank empty-shard fix, the new bench line at full config-2 size.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "rowpass" > gpurun_out/r2_03_rowpass_tests.log 2>&1
echo "rowpass tests exit $?"; tail -5 gpurun_out/r2_03_rowpass_tests.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider -k "many_rank" > gpurun_out/r2_03_multi.log 2>&1
echo "many-rank exit $?"; grep "multi-rank parity\|passed\|failed\|Error" gpurun_out/r2_03_multi.log | tail -8
timeout 600 python -m pytest tests/test_gpu_aoadmm.py tests/test_gpu_baseline_widths.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_03_e2e_tests.log 2>&1
echo "aoadmm + widths exit $?"; tail -3 gpurun_out/r2_03_e2e_tests.log
timeout 600 python tools/ab_rowpass.py 4096 > gpurun_out/r2_03_ab_rowpass.log 2>&1
echo "ab exit $?"; tail -7 gpurun_out/r2_03_ab_rowpass.log
timeout 900 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r2_03_bench_c2.json 2> gpurun_out/r2_03_bench_c2.err
echo "bench c2 exit $?"; cut -c1-400 gpurun_out/r2_03_bench_c2.json; tail -3 gpurun_out/r2_03_bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pf2_rowpass" -s 10 -c 3 -o gpurun_out/r2_03_prof_rowpass \
   python bench.py --config c2 --slices 2048 --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_03_ncu_rowpass.log 2>&1
echo "ncu rowpass exit $?"; tail -2 gpurun_out/r2_03_ncu_rowpass.log
