#!/bin/bash
# Evidence run v9: full -m gpu suite, smoke, benches of all configs (+ fp32 config 3, + reference arm), launch list of c2/c3, phase times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 exit $?"; cut -c1-260 gpurun_out/bench_c2.json
for c in c1 c3 c3f32 c4; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"; cut -c1-200 gpurun_out/bench_$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "ref exit $?"
timeout 300 python tools/phase_times.py > gpurun_out/phase_c2.log 2>&1; head -9 gpurun_out/phase_c2.log
timeout 300 python tools/phase_times.py c3 > gpurun_out/phase_c3.log 2>&1; head -7 gpurun_out/phase_c3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2s.csv python bench.py --config c2 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c2s.log 2>&1; echo "ncu c2s exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3s.csv python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3s.log 2>&1; echo "ncu c3s exit $?"
