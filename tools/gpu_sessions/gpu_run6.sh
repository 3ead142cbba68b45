#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unimodal_kernel" -s 6 -c 1 -o gpurun_out/prof_unimodal_c3 python bench.py --config c3 --slices 1024 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_uni.log 2>&1; echo "ncu uni exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pf2_(rowpass_mma|polar_cta|apply)_kernel" -s 18 -c 3 -o gpurun_out/prof_pf2_c2 python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_pf2.log 2>&1; echo "ncu pf2 exit $?"
ls -la gpurun_out/*.ncu-rep
