"""X-stream kernel rates (CUDA events) at the BASELINE shapes + fp64 pipe peaks. Writes gpurun_out/probe_xstream.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

out = {}
torch.cuda.set_device(0)


def ev_time(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for kind, name in ((0, "fp64_fma"), (1, "dmma_8x8x4"), (3, "dmma_plus_dfma")):
    best = 0
    for _ in range(3):
        flops, ms = _ops.microbench_flops(kind, 20000)
        best = max(best, flops / ms / 1e9)
    out[f"peak_{name}_tflops"] = best
    print(name, "TFLOP/s", best, flush=True)

shapes = ((4096 * 256, 512, 16, "c1"), (2048 * 1152, 1024, 20, "c2"), (1024 * 512, 2048, 32, "c4"),
          (4096 * 1024, 256, 8, "c3"), (2048 * 1152, 1024, 24, "r24"), (2048 * 1152, 1024, 12, "r12"))
for (N, K, R, tag) in shapes:
    X = torch.empty((N, K), dtype=torch.float64, device="cuda").normal_()
    C = torch.rand((K, R), dtype=torch.float64, device="cuda")
    W = _ops.alloc_w(N, R, torch.float64, "cuda", _lib.VARIANT_DMMA)
    W[:N, :R] = torch.rand((N, R), dtype=torch.float64, device="cuda")
    Y = torch.empty((N, R), dtype=torch.float64, device="cuda")
    Z = torch.empty((K, R), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    gb = N * K * 8 / 1e9
    ty = ev_time(lambda: _ops.xstream_y(X, N, K, C, Y, ws, _lib.VARIANT_DMMA))
    tz = ev_time(lambda: _ops.xstream_z(X, N, K, W, Z, ws, _lib.VARIANT_DMMA))
    n_chk = min(N, 1 << 16)
    ey = float((Y[:n_chk] - X[:n_chk] @ C).abs().max() / (X[:n_chk] @ C).abs().max())
    Zr = X.T @ W[:N, :R]
    ez = float((Z - Zr).abs().max() / Zr.abs().max())
    out[tag] = dict(N=N, K=K, R=R, y_ms=ty, z_ms=tz, y_gbs=gb / ty * 1e3, z_gbs=gb / tz * 1e3, y_err=ey, z_err=ez)
    print(tag, out[tag], flush=True)
    del X, C, W, Y, Z, ws, Zr
    torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_xstream.json"), "w"), indent=1)
