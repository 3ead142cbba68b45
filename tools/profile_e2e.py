"""Where does the host-facing cmf_aoadmm call spend its time? (bench.py's e2e leg on the c2 sample)"""
import cProfile, io, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matcouply_b200 import cmf_aoadmm  # noqa: E402
cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = dict(bench.CONFIGS[cfgname])
sizes = bench.slice_sizes(cfg)
S = bench.cpu_sample_size(cfg)
mats = bench.gen_host_sample(cfg, sizes, S)
kw = dict(cfg["kw"], random_state=0, tol=None, absolute_tol=None)
cmf_aoadmm(mats[:4], cfg["R"], n_iter_max=1, **kw)
torch.cuda.synchronize()
for k in (5, 50):
    t0 = time.perf_counter()
    cmf_aoadmm(mats, cfg["R"], n_iter_max=k, return_errors=True, **kw)
    torch.cuda.synchronize()
    print(f"{cfgname}: {S} slices, n_iter_max={k}: {time.perf_counter() - t0:.3f} s", flush=True)
pr = cProfile.Profile()
pr.enable()
cmf_aoadmm(mats, cfg["R"], n_iter_max=5, return_errors=True, **kw)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())
