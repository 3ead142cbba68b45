"""Stress the X-stream kernels for timing-dependent errors: repeat Y / Z many times, compare every run to a reference."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

torch.cuda.set_device(0)
for (N, K, R) in ((2359296, 1024, 20), (1048576, 512, 16), (524288, 2048, 32)):
    X = torch.empty((N, K), dtype=torch.float64, device="cuda").normal_()
    C = torch.rand((K, R), dtype=torch.float64, device="cuda")
    W = _ops.alloc_w(N, R, torch.float64, "cuda", _lib.VARIANT_DMMA)
    W[:N, :R] = torch.rand((N, R), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    Yref = torch.empty((N, R), dtype=torch.float64, device="cuda")
    for r0 in range(0, N, 1 << 18):
        Yref[r0:r0 + (1 << 18)] = X[r0:r0 + (1 << 18)] @ C
    Zref = X.T @ W[:N, :R]
    scale_y, scale_z = float(Yref.abs().max()), float(Zref.abs().max())
    bad_y = bad_z = 0
    for it in range(40):
        Y = torch.full((N, R), float("nan"), dtype=torch.float64, device="cuda")
        Z = torch.full((K, R), float("nan"), dtype=torch.float64, device="cuda")
        _ops.xstream_y(X, N, K, C, Y, ws, _lib.VARIANT_DMMA)
        if it % 2:
            _ops.xstream_z(X, N, K, W, Z, ws, _lib.VARIANT_DMMA)
        torch.cuda.synchronize()
        err = (Y - Yref).abs()
        e = float(torch.nan_to_num(err, nan=1e30).max()) / scale_y
        if e > 1e-12:
            bad_y += 1
            idx = torch.nonzero(torch.nan_to_num(err, nan=1e30) > 1e-9 * scale_y)
            rows = idx[:, 0].cpu().numpy()
            cols = idx[:, 1].cpu().numpy()
            print(f"Y run {it}: err {e:.3e}, {len(rows)} bad elements, rows {rows[:8]} (tile {rows[:8] // 256}, in-tile {rows[:8] % 256}) "
                  f"cols {np.unique(cols)[:24]}, n_tiles_bad {len(np.unique(rows // 256))}, row range {rows.min()}..{rows.max()}", flush=True)
        if it % 2:
            ez = float(torch.nan_to_num((Z - Zref).abs(), nan=1e30).max()) / scale_z
            if ez > 1e-11:
                bad_z += 1
                idx = torch.nonzero(torch.nan_to_num((Z - Zref).abs(), nan=1e30) > 1e-9 * scale_z)
                print(f"Z run {it}: err {ez:.3e}, {idx.shape[0]} bad, k {idx[:8, 0].cpu().numpy()} cols {idx[:8, 1].cpu().numpy()}", flush=True)
    print(f"shape {(N, K, R)}: bad Y runs {bad_y}/40, bad Z runs {bad_z}/20", flush=True)
    del X, C, W, Yref, Zref, ws
    torch.cuda.empty_cache()
