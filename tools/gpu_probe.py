"""Round-1 measurement probe (run on the B200 box): fp64 FMA / DMMA / fp32 peaks, copy bandwidth, X-stream kernel
rates for both fp64 variants at BASELINE config-1 size, and a first outer-iteration timing. Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

out = {}
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), flush=True)


def ev_time(fn, reps=5, warm=2):
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
    return float(np.median(ts)), float(np.min(ts))


for kind, name in ((0, "fp64_fma"), (1, "dmma_8x8x4"), (2, "fp32_fma")):
    best = 0
    for _ in range(3):
        flops, ms = _ops.microbench_flops(kind, 20000)
        best = max(best, flops / ms / 1e9)
    out[f"peak_{name}_tflops"] = best
    print(name, "TFLOP/s", best, flush=True)

# copy bandwidth (read+write), fp64 1 GiB
a = torch.empty(1 << 27, dtype=torch.float64, device="cuda").normal_()
b = torch.empty_like(a)
med, mn = ev_time(lambda: b.copy_(a))
out["copy_gbs"] = 2 * a.numel() * 8 / mn / 1e6
print("copy GB/s", out["copy_gbs"], flush=True)
med, mn = ev_time(lambda: a.sum())
out["read_gbs_torch_sum"] = a.numel() * 8 / mn / 1e6
print("torch.sum read GB/s", out["read_gbs_torch_sum"], flush=True)
del a, b

for (N, K, R, tag) in ((4096 * 256, 512, 16, "c1"), (2048 * 1024, 1024, 20, "c2slice"), (1024 * 512, 2048, 32, "c4slice"),
                       (4096 * 512, 256, 8, "c3slice")):
    X = torch.empty((N, K), dtype=torch.float64, device="cuda").normal_()
    C = torch.rand((K, R), dtype=torch.float64, device="cuda")
    Wv = {v: _ops.alloc_w(N, R, torch.float64, "cuda", v) for v in (_lib.VARIANT_FMA, _lib.VARIANT_DMMA)}
    for v in Wv:
        Wv[v][:N, :R] = torch.rand((N, R), dtype=torch.float64, device="cuda")
    Y = torch.empty((N, R), dtype=torch.float64, device="cuda")
    Z = torch.empty((K, R), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    xb = N * K * 8
    for variant, vn in ((_lib.VARIANT_FMA, "fma"), (_lib.VARIANT_DMMA, "dmma")):
        med, mn = ev_time(lambda: _ops.xstream_y(X, N, K, C, Y, ws, variant))
        out[f"y_{tag}_{vn}_gbs"] = xb / med / 1e6
        print(f"xstream_y {tag} {vn}: {med:.3f} ms  {xb / med / 1e6:.0f} GB/s", flush=True)
    for variant, vn in ((_lib.VARIANT_FMA, "fma"), (_lib.VARIANT_DMMA, "dmma")):
        med, mn = ev_time(lambda: _ops.xstream_z(X, N, K, Wv[variant], Z, ws, variant))
        out[f"z_{tag}_{vn}_gbs"] = xb / med / 1e6
        print(f"xstream_z {tag} {vn}: {med:.3f} ms  {xb / med / 1e6:.0f} GB/s", flush=True)
    o = torch.zeros(1, dtype=torch.float64, device="cuda")
    med, mn = ev_time(lambda: _ops.sumsq(X, N, K, o, ws))
    out[f"sumsq_{tag}_gbs"] = xb / med / 1e6
    print(f"sumsq {tag}: {med:.3f} ms {xb / med / 1e6:.0f} GB/s", flush=True)
    if tag == "c3slice":
        X32 = X.float()
        ws32 = _ops.Workspace("cuda", K, R, torch.float32)
        Y32 = torch.empty((N, R), dtype=torch.float32, device="cuda")
        Z32 = torch.empty((K, R), dtype=torch.float32, device="cuda")
        med, _ = ev_time(lambda: _ops.xstream_y(X32, N, K, C.float(), Y32, ws32, _lib.VARIANT_FMA))
        print(f"xstream_y f32 {tag}: {med:.3f} ms {xb / 2 / med / 1e6:.0f} GB/s", flush=True)
        out[f"y_{tag}_f32_gbs"] = xb / 2 / med / 1e6
        W32 = _ops.alloc_w(N, R, torch.float32, "cuda", _lib.VARIANT_FMA)
        W32[:N, :R] = torch.rand((N, R), dtype=torch.float32, device="cuda")
        med, _ = ev_time(lambda: _ops.xstream_z(X32, N, K, W32, Z32, ws32, _lib.VARIANT_FMA))
        print(f"xstream_z f32 {tag}: {med:.3f} ms {xb / 2 / med / 1e6:.0f} GB/s", flush=True)
        out[f"z_{tag}_f32_gbs"] = xb / 2 / med / 1e6
        del X32
    del X, C, Wv, Y, Z

# outer iteration at config 1 (NN-CMF) through the engine, device-generated data
from matcouply_b200._engine import AOADMMEngine, PackedMatrices  # noqa: E402
from matcouply_b200 import penalties as P  # noqa: E402

for tag, I, K, J, R, regs_fn in (
    ("c1_nn", 4096, 512, 256, 16, lambda: [[P.NonNegativity()], [P.NonNegativity()], [P.NonNegativity()]]),
    ("pf2_nn_l1", 2048, 1024, 512, 20, lambda: [[P.NonNegativity()], [P.Parafac2(), P.NonNegativity()],
                                                  [P.L1Penalty(0.1, non_negativity=True)]]),
    ("c3_uni", 2048, 256, 1024, 8, lambda: [[P.NonNegativity()], [P.Parafac2(), P.Unimodality(True), P.L2Ball(1.0, True)],
                                             [P.L2Ball(1.0, True)]]),
):
    N = I * J
    X = torch.rand((N, K), dtype=torch.float64, device="cuda")
    packed = PackedMatrices(X, np.arange(I + 1, dtype=np.int64) * J, K)
    regs = regs_fn()
    eng = AOADMMEngine(packed, R, regs)
    rs = np.random.RandomState(0)
    mats = [type("S", (), {"shape": (J, K)})() for _ in range(I)]
    A0, C0, B0 = rs.uniform(size=(I, R)), rs.uniform(size=(K, R)), rs.uniform(size=(N, R))
    aux = [[r.init_aux(mats, R, m, rs) for r in regs[m]] for m in range(3)]
    dual = [[r.init_dual(mats, R, m, rs) for r in regs[m]] for m in range(3)]
    eng.load_state(A0, B0, C0, aux, dual)
    eng.prepare()
    t0 = time.time()
    med, mn = ev_time(lambda: (eng.outer_iteration(), eng.diagnostics()), reps=5, warm=2)
    out[f"iter_{tag}_ms"] = med
    print(f"outer iteration {tag}: {med:.2f} ms  -> {1000 / med:.1f} it/s ; X={N * K * 8 / 1e9:.2f} GB "
          f"eff stream {2 * N * K * 8 / med / 1e6:.0f} GB/s", flush=True)
    # per-phase timing
    for nm, fn in (("step_B", eng.step_B), ("step_C", eng.step_C), ("refresh", eng.refresh_products),
                   ("step_A", eng.step_A), ("diag", eng.diagnostics)):
        m2, _ = ev_time(fn, reps=3, warm=1)
        out[f"{tag}_{nm}_ms"] = m2
        print(f"   {nm}: {m2:.3f} ms", flush=True)
    del eng, X, packed

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print(json.dumps(out))
