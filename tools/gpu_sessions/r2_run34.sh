#!/bin/bash
# Round 2, GPU session 34 (1 GPU): ncu --set full with source of the shipped unimodal kernel (variant 9) at 2 048 slices
# of config 3 and of the polar kernel at 2 048 slices of config 2 (stall attribution per source line).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"unimodal_kernel" -s 2 -c 1 -o gpurun_out/r2_34_prof_unimodal \
   python bench.py --config c3 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_34_ncu_unimodal.log 2>&1
echo "ncu unimodal exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pf2_polar_reg" -s 7 -c 1 -o gpurun_out/r2_34_prof_polar \
   python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_34_ncu_polar.log 2>&1
echo "ncu polar exit $?"
ls -la gpurun_out/r2_34_*
