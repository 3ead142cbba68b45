"""Do the DMMA and DFMA pipes overlap on B200?  Times microbench kinds 0 (DFMA), 1 (DMMA), 3 (8 DMMA + 16 DFMA interleaved)
with the same iteration count.  If t3 ~= t1 the pipes are independent; if t3 ~= t1 + t0 * (1024/8192) ... they share."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _ops  # noqa: E402
torch.cuda.set_device(0)
out = {}
for kind in (0, 1, 3, 0, 1, 3):
    best = None
    for _ in range(3):
        flops, ms = _ops.microbench_flops(kind, 20000)
        best = ms if best is None else min(best, ms)
    out[f"kind{kind}_ms"] = best
    out[f"kind{kind}_tflops"] = flops / best / 1e9
    print(kind, best, flops / best / 1e9, flush=True)
# per-warp per-iteration: kind0 = 16 DFMA (x8 warps x 4 CTAs/SM), kind1 = 8 DMMA, kind3 = both
out["t3_over_t1"] = out["kind3_ms"] / out["kind1_ms"]
out["t0_over_t1"] = out["kind0_ms"] / out["kind1_ms"]
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_pipes.json"), "w"), indent=1)
