#!/bin/bash
# Round 2, GPU session 50 (1 GPU): ncu --set full of the shipped unimodal kernels (variant 14: PAVA kernel + fill kernel)
# at full config-3 size: DRAM traffic per launch for profiles/traffic.json.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"unimodal" -s 4 -c 2 -o gpurun_out/r2_50_prof_unimodal_v14_full \
   python tools/prof_unimodal.py 14 noise 4 > gpurun_out/r2_50_ncu_unimodal.log 2>&1
echo "ncu unimodal exit $?"; tail -2 gpurun_out/r2_50_ncu_unimodal.log
