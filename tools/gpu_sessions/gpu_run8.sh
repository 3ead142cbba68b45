#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pf2_(rowpass_mma|polar_cta)_kernel" -s 14 -c 2 -o gpurun_out/prof_pf2b_c2 python bench.py --config c2 --slices 2048 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_pf2b.log 2>&1; echo "ncu pf2 exit $?"
timeout 800 ncu --set full --clock-control none --import-source on -k regex:"unimodal_kernel" -s 6 -c 1 -o gpurun_out/prof_unimodal4_c3full python bench.py --config c3 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_uni4.log 2>&1; echo "ncu uni exit $?"
