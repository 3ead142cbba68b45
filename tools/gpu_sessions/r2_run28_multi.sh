#!/bin/bash
# Round 2, multi-GPU re-check after the last engine changes: `gpurun --gpus 2 -- bash tools/gpu_sessions/r2_run28_multi.sh 2`
# NCCL parity tests on N GPUs, the headline config and the CMF diagnostic config (fused single-read pass, sharded) at N ranks,
# the reference arm under torchrun (rank 0 only).
N=${1:-2}
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider -k "not many_rank" > gpurun_out/r2_28_multi_nccl_n$N.log 2>&1
echo "multi tests exit $?"; grep "passed\|failed\|Error" gpurun_out/r2_28_multi_nccl_n$N.log | tail -5
for c in c2 c1r8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --config $c > gpurun_out/r2_28_bench_${c}_n$N.json 2> gpurun_out/r2_28_bench_${c}_n$N.err
echo "$c N=$N exit $?"; tail -2 gpurun_out/r2_28_bench_${c}_n$N.err; cut -c1-400 gpurun_out/r2_28_bench_${c}_n$N.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/r2_28_bench_reference_n$N.json 2> gpurun_out/r2_28_bench_reference_n$N.err
echo "reference arm N=$N exit $?"; cut -c1-300 gpurun_out/r2_28_bench_reference_n$N.json
