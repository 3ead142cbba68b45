#!/bin/bash
# Round 2, 2-GPU re-check of the final state: `gpurun --gpus 2 -- bash tools/gpu_sessions/r2_run46_multi.sh 2`
# NCCL parity tests on N GPUs (incl. the sharded unimodal / L2Ball cases), config 3 (new unimodal kernels) and the
# headline config at N ranks.
N=${1:-2}
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider -k "not many_rank" > gpurun_out/r2_46_multi_nccl_n$N.log 2>&1
echo "multi tests exit $?"; grep "passed\|failed\|Error" gpurun_out/r2_46_multi_nccl_n$N.log | tail -5
for c in c3 c2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --config $c --no-cpu > gpurun_out/r2_46_bench_${c}_n$N.json 2> gpurun_out/r2_46_bench_${c}_n$N.err
echo "$c N=$N exit $?"; tail -2 gpurun_out/r2_46_bench_${c}_n$N.err; cut -c1-300 gpurun_out/r2_46_bench_${c}_n$N.json
done
