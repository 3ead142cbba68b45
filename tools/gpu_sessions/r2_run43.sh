#!/bin/bash
# Round 2, GPU session 43 (1 GPU): Jacobi warm start across B-updates (cold every 8th): PARAFAC2 trajectory parity,
# A/B of the polar step at 4 096 slices of config 2 (B2_POLAR_COLD_EVERY=1 is the previous behaviour).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_aoadmm.py tests/test_gpu_baseline_widths.py tests/test_gpu_penalty_contract.py tests/test_gpu_reference_kats.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_43_tests_aoadmm.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/r2_43_tests_aoadmm.log
for ce in 1 8; do
  B2_POLAR_COLD_EVERY=$ce timeout 600 python bench.py --config c2 --slices 4096 --steps 16 --warmup 4 --no-cpu > gpurun_out/r2_43_bench_c2s_cold$ce.json 2> gpurun_out/r2_43_bench_c2s_cold$ce.err
  echo "bench cold_every=$ce exit $?"; python - <<P
import json
d=json.loads(open("gpurun_out/r2_43_bench_c2s_cold$ce.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k: round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
P
done
