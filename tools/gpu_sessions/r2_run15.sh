#!/bin/bash
# Round 2, GPU session 15 (1 GPU): full GPU suite after the fused single-read pass, the unimodal kernel variant 9 and the
# cached C^T C; bench lines c1 / c1r8 / c3 / c2; ncu of the unimodal kernel (traffic) and of the fused pass at R = 8.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_15_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_15_tests.log
for c in c1 c1r8 c3; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_15_bench_$c.json 2> gpurun_out/r2_15_bench_$c.err
  echo "bench $c exit $?"; cut -c1-200 gpurun_out/r2_15_bench_$c.json
done
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --x1 on --no-cpu > gpurun_out/r2_15_bench_c1_x1on.json 2> gpurun_out/r2_15_bench_c1_x1on.err
echo "bench c1 x1=on exit $?"; cut -c1-200 gpurun_out/r2_15_bench_c1_x1on.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"unimodal_kernel" -s 2 -c 1 -o gpurun_out/r2_15_prof_unimodal \
   python bench.py --config c3 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_15_ncu_unimodal.log 2>&1
echo "ncu unimodal exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"xfused_local" -s 2 -c 1 -o gpurun_out/r2_15_prof_xfused_c1r8 \
   python bench.py --config c1r8 --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_15_ncu_xfused_r8.log 2>&1
echo "ncu xfused r8 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_15_launches_c1r8.csv \
   python bench.py --config c1r8 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_15_launches_c1r8.log 2>&1
echo "launch list exit $?"
timeout 1500 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r2_15_bench_c2.json 2> gpurun_out/r2_15_bench_c2.err
echo "bench c2 exit $?"; cut -c1-200 gpurun_out/r2_15_bench_c2.json
