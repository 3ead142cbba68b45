"""Per-kernel-family device time of one outer iteration at full bench size (CUDA events around every _ops call)."""
import collections, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matcouply_b200 import _ops  # noqa: E402
from matcouply_b200._engine import AOADMMEngine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
slices = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = dict(bench.CONFIGS[name])
if slices:
    cfg["I"] = slices
sizes = bench.slice_sizes(cfg)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dtype = torch.float64
packed = bench.gen_device_data(cfg, sizes, 0, cfg["I"], dtype, dev)
eng = AOADMMEngine(packed, cfg["R"], bench.make_regs(cfg["kw"]))
eng.load_state_device(seed=0)
eng.prepare()
for _ in range(2):
    eng.outer_iteration(); eng.diagnostics()
torch.cuda.synchronize()
acc = collections.OrderedDict()
orig = {}
def wrap(fname):
    f = getattr(_ops, fname)
    orig[fname] = f
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record(); torch.cuda.synchronize()
        acc.setdefault(fname, []).append(e0.elapsed_time(e1)); return r
    setattr(_ops, fname, g)
for fn in ["xstream_y", "xstream_z", "gram", "scale_gram", "rho_from_trace", "factor_batch", "admm_local", "admm_solve",
           "pf2_rowpass", "pf2_polar", "pf2_delta", "pf2_apply", "pf2_gap", "prox_l2ball", "prox_unimodal", "slice_gram",
           "slice_coldot", "weighted_gram_sum", "hadamard_bcast", "reduce_stats", "fit_terms", "rowscale", "slice_cross"]:
    if hasattr(_ops, fn):
        wrap(fn)
n = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(n):
    eng.outer_iteration(); eng.diagnostics()
tot = sum(sum(v) for v in acc.values()) / n
out = {"config": name, "slices": int(cfg["I"]), "rows": int(packed.N), "sum_ms_per_iter": tot, "ops": {}}
print(f"{name}: {cfg['I']} slices, {packed.N} rows; sum of op times per outer iteration {tot:.3f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    out["ops"][k] = {"calls_per_iter": len(v) / n, "ms_per_iter": sum(v) / n, "ms_per_call": float(np.mean(v))}
    print(f"  {sum(v)/n:9.3f} ms {100*sum(v)/n/tot:5.1f}%  x{len(v)/n:5.1f}  avg {np.mean(v):8.3f} ms  {k}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"phase_{name}.json"), "w"), indent=1)
