#!/bin/bash
# Round-1 evidence run: full -m gpu suite, smoke, benches of all configs, launch lists, full ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -2 gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_c2_ref.json
for c in c1 c3 c4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"; cut -c1-330 gpurun_out/bench_$c.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2s.csv python bench.py --config c2 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c2s.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3s.csv python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3s.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --config c1 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c1.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xstream_(y|z_dmma)_kernel|pf2_(rowpass_mma|polar_cta)_kernel" -s 14 -c 6 -o gpurun_out/prof_c2_top python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unimodal_kernel|admm_local_mma_kernel" -s 4 -c 2 -o gpurun_out/prof_c3_top python bench.py --config c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full3.log 2>&1; echo "ncu full3 exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3
