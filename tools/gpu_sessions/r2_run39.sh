#!/bin/bash
# Round 2, GPU session 39 (1 GPU): ncu --set full with source of the compact-record unimodal kernel (variant 16) at full config-3 size.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"unimodal2_kernel" -s 2 -c 1 -o gpurun_out/r2_39_prof_unimodal_v16_full \
   python tools/prof_unimodal.py 16 noise 4 > gpurun_out/r2_39_ncu_unimodal.log 2>&1
echo "ncu unimodal exit $?"; tail -2 gpurun_out/r2_39_ncu_unimodal.log
