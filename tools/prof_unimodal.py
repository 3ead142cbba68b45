"""One unimodal-kernel variant at config-3 size on noise-like or peak-like input, a few launches (for ncu):
    python tools/prof_unimodal.py VARIANT [noise|peaks] [reps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

variant = int(sys.argv[1]); kind = sys.argv[2] if len(sys.argv) > 2 else "noise"; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lib = _lib.load()
lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, variant)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
G, J, R = 8192, 1024, 8
off = torch.arange(0, (G + 1) * J, J, dtype=torch.int64, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
dtype = torch.float64
if kind == "noise":
    V = torch.randn((G * J, R), dtype=dtype, device=dev, generator=gen)
else:
    t = torch.arange(J, dtype=dtype, device=dev)[None, :, None]
    centre = torch.rand((G, 1, R), dtype=dtype, device=dev, generator=gen) * 0.6 * J + 0.2 * J
    width = torch.rand((G, 1, R), dtype=dtype, device=dev, generator=gen) * 0.1 * J + 0.03 * J
    V = (torch.exp(-0.5 * ((t - centre) / width) ** 2) + 0.02 * torch.randn((G, J, R), dtype=dtype, device=dev, generator=gen)).reshape(G * J, R).contiguous()
ws = _ops.Workspace(dev, 256, R, dtype, unimodal_shape=(G, R, J))
times = []
for rep in range(reps):
    aux, dual = torch.empty_like(V), V.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _ops.prox_unimodal(aux, dual, off, G, R, J, True, ws); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
print({"variant": variant, "input": kind, "ms": times})
