"""Y-kernel stress at the c2 shape (flaky-error hunt). Usage: stress_y2.py [runs]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.cuda.set_device(0)
N, K, R = 2359296, 1024, 20
X = torch.empty((N, K), dtype=torch.float64, device="cuda").normal_()
C = torch.rand((K, R), dtype=torch.float64, device="cuda")
ws = _ops.Workspace("cuda", K, R, torch.float64)
Yref = torch.empty((N, R), dtype=torch.float64, device="cuda")
for r0 in range(0, N, 1 << 18):
    Yref[r0:r0 + (1 << 18)] = X[r0:r0 + (1 << 18)] @ C
scale = float(Yref.abs().max())
bad = 0
Y = torch.empty((N, R), dtype=torch.float64, device="cuda")
for it in range(runs):
    Y.fill_(float("nan"))
    _ops.xstream_y(X, N, K, C, Y, ws, _lib.VARIANT_DMMA)
    torch.cuda.synchronize()
    err = torch.nan_to_num((Y - Yref).abs(), nan=1e30)
    e = float(err.max()) / scale
    if e > 1e-12:
        bad += 1
        rows = torch.nonzero(err.max(dim=1).values > 1e-9 * scale)[:, 0].cpu().numpy()
        print(f"run {it}: err {e:.2e} bad rows {len(rows)}: {rows[:12]} tiles128 {np.unique(rows // 128)[:6]}", flush=True)
print(f"LIB={os.environ.get('B2_LIB_PATH_DEBUG','new')} Y_DEBUG={os.environ.get('B2_Y_DEBUG','')}: bad runs {bad}/{runs}", flush=True)
