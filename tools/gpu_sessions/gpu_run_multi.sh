#!/bin/bash
# N-GPU validation: NCCL trajectory-parity tests + the sharded bench (strong scaling of config 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_multi.log 2>&1
echo "multi tests exit $?" >> gpurun_out/tests_multi.log; tail -8 gpurun_out/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench N=$N exit $?"; tail -5 gpurun_out/bench_c2_n$N.err; cat gpurun_out/bench_c2_n$N.json
