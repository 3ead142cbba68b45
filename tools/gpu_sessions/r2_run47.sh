#!/bin/bash
# Round 2, GPU session 47 (1 GPU): A/B of the page-locked upload over 1-4 streams (24 GB).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python tools/ab_upload.py 24 > gpurun_out/r2_47_ab_upload.log 2>&1
echo "exit $?"; cat gpurun_out/r2_47_ab_upload.log
