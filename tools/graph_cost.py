import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from matcouply_b200 import cmf_aoadmm
cfg = dict(bench.CONFIGS["c1"]); cfg["I"] = 1024
sizes = bench.slice_sizes(cfg)
X = bench.gen_host_sample(cfg, sizes, cfg["I"])
for mode in (False, True, True, True, False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cmf_aoadmm(X, cfg["R"], n_iter_max=50, tol=None, absolute_tol=None, random_state=0, return_errors=True, use_cuda_graph=mode, **cfg["kw"])
    torch.cuda.synchronize(); print("graph", mode, round(time.perf_counter() - t0, 4), "s", flush=True)
