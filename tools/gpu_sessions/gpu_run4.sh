#!/bin/bash
# HEAD validation: full -m gpu suite, smoke, benches (c2 with cpu legs), launch lists, full ncu capture of the X-stream kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -25 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 600 python bench.py --config c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 exit $?"; tail -3 gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 exit $?"; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2s.csv python bench.py --config c2 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c2s.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3s.csv python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3s.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --config c1 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c1.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xstream_(y|z_dmma)_kernel" -s 2 -c 4 -o gpurun_out/prof_xstream_c2 python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
