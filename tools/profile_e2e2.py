"""Phase timing of the host-facing call on the bench's e2e sample (page-locked input)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matcouply_b200 import decomposition as D, penalties  # noqa: E402
from matcouply_b200._engine import AOADMMEngine, PackedMatrices  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cfg = dict(bench.CONFIGS[name]); sizes = bench.slice_sizes(cfg)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
S2 = bench.e2e_sample_size(cfg, 8)
views, rows = bench.gen_pinned_sample(dict(cfg, I=S2), sizes[:S2], 0, S2, torch.float64, dev)
kw = dict(cfg["kw"], random_state=0, tol=None, absolute_tol=None)
D.cmf_aoadmm(views, cfg["R"], n_iter_max=1, **kw); torch.cuda.synchronize()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t = [T()]
    rs = np.random.RandomState(0)
    packed = PackedMatrices.from_list(views, torch.float64, dev); t.append(T())
    A0, B0, C0 = D.initialize_cmf(views, cfg["R"], "random", random_state=rs, _device=dev); t.append(T())
    regs = bench.make_regs(cfg["kw"])
    penalties._DEVICE_DRAW["device"] = dev
    auxes = [[r.init_aux(views, cfg["R"], m, random_state=rs) for r in regs[m]] for m in range(3)]
    duals = [[r.init_dual(views, cfg["R"], m, random_state=rs) for r in regs[m]] for m in range(3)]; t.append(T())
    penalties._DEVICE_DRAW["device"] = None
    eng = AOADMMEngine(packed, cfg["R"], regs); t.append(T())
    eng.load_state(np.asarray(A0), B0, np.asarray(C0), auxes, duals); t.append(T())
    eng.prepare(); eng.diagnostics(); t.append(T())
    for it in range(50):
        eng.outer_iteration(); eng.diagnostics()
    t.append(T())
    f = eng.factors(); t.append(T())
    names = ["from_list(H2D X)", "initialize_cmf (RNG)", "init aux/dual (RNG)", "engine ctor", "load_state (H2D state)", "prepare+diag", "50 iterations", "factors (D2H)"]
    print(f"{name}: {S2} slices {rows} rows")
    for n, a, b in zip(names, t[:-1], t[1:]):
        print(f"   {n:28s} {1000*(b-a):9.1f} ms")
    print(f"   total {1000*(t[-1]-t[0]):.1f} ms")
    t0 = T(); D.cmf_aoadmm(views, cfg["R"], n_iter_max=50, return_errors=True, **kw); t1 = T()
    print(f"   cmf_aoadmm(n_iter_max=50) whole call: {1000*(t1-t0):.1f} ms")
    del eng, packed
