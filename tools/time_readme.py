"""Latency of the README example (BASELINE config 0: 15 slices 50x20, R=3, five penalties) through cmf_aoadmm."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import cmf_aoadmm  # noqa: E402
g = np.load(os.path.join(ROOT, "tests", "golden", "traj_c0_readme.npz"), allow_pickle=False)
off = g["row_offsets"]
X = [g["X"][a:b] for a, b in zip(off[:-1], off[1:])]
kw = json.loads(str(g["kwargs"]))
for key in ("l1_penalty", "non_negative", "unimodal", "l2_norm_bound"):
    if isinstance(kw.get(key), dict):
        kw[key] = {int(k): v for k, v in kw[key].items()}
kw.pop("n_iter_max", None); kw.pop("tol", None); kw.pop("absolute_tol", None)
R = int(g["rank"])
cmf_aoadmm(X, R, n_iter_max=3, **kw)
torch.cuda.synchronize()
for n in (100, 1000):
    t0 = time.perf_counter()
    out = cmf_aoadmm(X, R, n_iter_max=n, return_errors=True, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"README config: n_iter_max={n}: ran {out[1].n_iter} iterations in {dt:.3f} s = {1000*dt/out[1].n_iter:.3f} ms/iteration; message: {out[1].message}", flush=True)
