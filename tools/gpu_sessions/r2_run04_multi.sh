#!/bin/bash
# Round 2, multi-GPU session: `gpurun --gpus N -- bash tools/gpu_sessions/r2_run04_multi.sh N [tests]`
# NCCL parity tests (2 / 4 / 8 ranks, one GPU per rank; device= keyword test), BASELINE config 4 as the weak-scaling
# sweep (8192 slices per GPU) and the strong-scaling headline config at N ranks.
N=${1:-8}
TESTS=${2:-yes}
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2_04_gpus_n$N.txt; free -g >> gpurun_out/r2_04_gpus_n$N.txt; nproc >> gpurun_out/r2_04_gpus_n$N.txt
if [ "$TESTS" = "yes" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_04_multi_nccl_n$N.log 2>&1
  echo "multi tests exit $?"; grep "multi-rank parity\|passed\|failed\|Error" gpurun_out/r2_04_multi_nccl_n$N.log | tail -25
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config c4w > gpurun_out/r2_04_bench_c4w_n$N.json 2> gpurun_out/r2_04_bench_c4w_n$N.err
echo "c4w N=$N exit $?"; tail -3 gpurun_out/r2_04_bench_c4w_n$N.err; cut -c1-700 gpurun_out/r2_04_bench_c4w_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_04_bench_c2_n$N.json 2> gpurun_out/r2_04_bench_c2_n$N.err
echo "c2 N=$N exit $?"; tail -3 gpurun_out/r2_04_bench_c2_n$N.err; cut -c1-700 gpurun_out/r2_04_bench_c2_n$N.json
