"""A/B of the fp64 X-stream kernel variants in one process, alternating (B2_OPT_XSTREAM_HYBRID: 0 = padded DMMA blocks,
8 consumer warps; 1 = DMMA blocks + DFMA remainder columns; 2 = Y kernel with 12 consumer warps / 384-row tiles).
    python tools/ab_xstream.py [rows] [rank] [variants, e.g. 0,2]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402
torch.cuda.set_device(0)
lib = _lib.load()
N, K, R = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 20
X = torch.empty((N, K), dtype=torch.float64, device="cuda").normal_()
C = torch.rand((K, R), dtype=torch.float64, device="cuda")
Y = torch.zeros((N, R), dtype=torch.float64, device="cuda")
Z = torch.zeros((K, R), dtype=torch.float64, device="cuda")
W = _ops.alloc_w(N, R, torch.float64, "cuda", _lib.VARIANT_DMMA)
W[:N, :R] = torch.rand((N, R), dtype=torch.float64, device="cuda")
ws = _ops.Workspace(torch.device("cuda"), K, R, torch.float64)
def t(fn, reps=6):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
VARIANTS = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (0, 1)
out = {}
for rnd in range(3):
    for hyb in VARIANTS:
        lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, hyb)
        ty = t(lambda: _ops.xstream_y(X, N, K, C, Y, ws, _lib.VARIANT_DMMA))
        if hyb == VARIANTS[0] and rnd == 0:
            Y_ref = Y.clone()
        elif rnd == 0:
            print(f"variant {hyb}: Y bit-identical to variant {VARIANTS[0]}: {bool(torch.equal(Y, Y_ref))}, max rel diff "
                  f"{float(((Y - Y_ref).abs().max() / Y_ref.abs().max()).item()):.2e}", flush=True)
        tz = t(lambda: _ops.xstream_z(X, N, K, W, Z, ws, _lib.VARIANT_DMMA))
        gb = N * K * 8 / 1e6
        print(f"round {rnd} hybrid={hyb}: Y {ty:.3f} ms ({gb/ty:.0f} GB/s)  Z {tz:.3f} ms ({gb/tz:.0f} GB/s)", flush=True)
        out.setdefault(f"hybrid{hyb}", []).append((ty, tz))
lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_xstream.json"), "w"))
