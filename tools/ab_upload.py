"""A/B of the host -> HBM upload of page-locked slices: one cudaMemcpyAsync on the current stream against chunks
round-robin over 2-4 side streams (PackedMatrices.UPLOAD_STREAMS / B2_UPLOAD_STREAMS).
    python tools/ab_upload.py [GB]  ->  gpurun_out/ab_upload.json"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200._engine import PackedMatrices  # noqa: E402

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 24.0
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
K, J = 1024, 1152
n_slices = int(gb * 1e9 / (K * J * 8))
host = torch.empty((n_slices * J, K), dtype=torch.float64, pin_memory=True)
host.normal_()
arr = host.numpy()
views = [arr[i * J:(i + 1) * J] for i in range(n_slices)]
out = {"gbytes": host.numel() * 8 / 1e9, "runs": []}
ref = None
for rep in range(2):
    for n_up in (1, 2, 3, 4):
        PackedMatrices.UPLOAD_STREAMS = n_up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = PackedMatrices.from_list(views, torch.float64, dev)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        chk = float(p.X[:, :K].sum().item())
        if ref is None:
            ref = chk
        rec = {"streams": n_up, "seconds": t, "gb_per_s": out["gbytes"] / t, "same_checksum": chk == ref}
        out["runs"].append(rec); print(rec)
        del p
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_upload.json"), "w"), indent=1)
