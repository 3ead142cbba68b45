#!/bin/bash
# Evidence run v8 (register polar kernel, new penalties): launch lists + full ncu captures + benches of all configs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2s.csv python bench.py --config c2 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c2s.log 2>&1; echo "ncu c2s exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3s.csv python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3s.log 2>&1; echo "ncu c3s exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --config c1 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c1.log 2>&1; echo "ncu c1 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xstream_(y|z_dmma)_kernel|pf2_(rowpass_mma|polar_reg)_kernel" -s 14 -c 6 -o gpurun_out/prof_c2_top python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"; cut -c1-200 gpurun_out/bench_$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_c2_ref.json
ls -la gpurun_out/*.ncu-rep | tail -3
