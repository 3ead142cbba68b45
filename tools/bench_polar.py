"""Device time of b2_pf2_polar at config-2 / config-3 size, warp-per-slice vs CTA-per-slice kernel (cold and warm)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

lib = _lib.load()
res = {}
for name, G, R in (("c2", 16384, 20), ("c3", 8192, 8), ("r16", 16384, 16), ("r24", 8192, 24)):
    gen = torch.Generator(device="cuda").manual_seed(0)
    V = torch.randn(G, 3 * R, R, dtype=torch.float64, device="cuda", generator=gen)
    S = torch.matmul(V.transpose(1, 2), V).contiguous()
    S2 = S + 1e-3 * torch.matmul(V.transpose(1, 2), torch.randn_like(V))
    S2 = (0.5 * (S2 + S2.transpose(1, 2))).contiguous()
    Delta = torch.rand(R, R, dtype=torch.float64, device="cuda", generator=gen) + torch.eye(R, dtype=torch.float64, device="cuda")
    rho = torch.rand(G, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    for variant in (2, 1, 0):
        lib.b2_set_option(_lib.OPT_POLAR_WARP, variant)
        Wm, num, Q = (torch.zeros(G, R, R, dtype=torch.float64, device="cuda") for _ in range(3))
        times = {"cold": [], "warm": []}
        for rep in range(4):
            for key, Sx, warm in (("cold", S, False), ("warm", S2, True)):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _ops.pf2_polar(Sx, Delta, rho, G, R, Wm, num, Q, warm=warm); e1.record()
                torch.cuda.synchronize()
                if rep:
                    times[key].append(e0.elapsed_time(e1))
        vname = ("cta", "warp", "reg")[variant]
        res[f"{name}_{vname}"] = {k: float(np.mean(v)) for k, v in times.items()}
        print(name, vname, res[f"{name}_{vname}"], flush=True)
lib.b2_set_option(_lib.OPT_POLAR_WARP, 2)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_polar.json"), "w"), indent=1)
