#!/bin/bash
# 2-GPU box: whole -m gpu suite (incl. the NCCL tests) + the c2 bench sharded over 2 ranks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_2gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_2gpu.log; tail -15 gpurun_out/tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "bench n2 exit $?"; tail -3 gpurun_out/bench_c2_n2.err; cut -c1-400 gpurun_out/bench_c2_n2.json
timeout 300 python tools/phase_times.py 2>&1 | head -6
