#!/bin/bash
# Round 2, GPU session 6: the single-read fused pass (csrc/xfused.cu) — engine-level parity (fused vs two-pass schedule,
# golden trajectories, BASELINE widths), config-1 bench with the fused pass and with the two-pass schedule (A/B),
# launch list + one `ncu --set full` capture of the fused kernel.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_aoadmm.py tests/test_gpu_baseline_widths.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_06_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_06_tests.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_reference_kats.py tests/test_gpu_penalty_contract.py -m gpu -q -x -p no:cacheprovider -k "not many_rank" > gpurun_out/r2_06_tests2.log 2>&1
echo "tests2 exit $?"; tail -3 gpurun_out/r2_06_tests2.log
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 > gpurun_out/r2_06_bench_c1_fused.json 2> gpurun_out/r2_06_bench_c1_fused.err
echo "bench c1 fused exit $?"; head -c 900 gpurun_out/r2_06_bench_c1_fused.json; echo
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --x1 off --no-cpu > gpurun_out/r2_06_bench_c1_twopass.json 2> gpurun_out/r2_06_bench_c1_twopass.err
echo "bench c1 two-pass exit $?"; head -c 600 gpurun_out/r2_06_bench_c1_twopass.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_06_launches_c1.csv \
   python bench.py --config c1 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_06_launches_c1.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"xfused_local" -s 2 -c 1 -o gpurun_out/r2_06_prof_xfused_c1 \
   python bench.py --config c1 --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_06_ncu_xfused.log 2>&1
echo "ncu xfused exit $?"; tail -2 gpurun_out/r2_06_ncu_xfused.log
