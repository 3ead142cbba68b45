"""A/B of the unimodal-regression kernel variants (B2_OPT_UNIMODAL_VARIANT: ring depth, stack-cache depth, CTAs per SM)
at config-3 size (8192 slices x 1024 rows x 8 columns, fp64 and fp32) on two kinds of input: noise-like pre-images (what
bench.py's synthetic config 3 feeds the kernel: PAVA stacks stay ~15 deep) and chromatography-like peaks (Gaussian
bumps + noise: the stack grows to hundreds of blocks on the rising flank).  Results must be bit-identical across variants.
    python tools/ab_unimodal.py  ->  gpurun_out/ab_unimodal.json"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matcouply_b200 import _lib, _ops  # noqa: E402

lib = _lib.load()
DEFAULT = lib.b2_get_option(_lib.OPT_UNIMODAL_VARIANT)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
G, J, R = 8192, 1024, 8
off = torch.arange(0, (G + 1) * J, J, dtype=torch.int64, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
out = {"G": G, "J": J, "R": R, "runs": []}
for dtype in (torch.float64, torch.float32):
    noise = torch.randn((G * J, R), dtype=dtype, device=dev, generator=gen)
    t = torch.arange(J, dtype=dtype, device=dev)[None, :, None]
    centre = torch.rand((G, 1, R), dtype=dtype, device=dev, generator=gen) * 0.6 * J + 0.2 * J
    width = torch.rand((G, 1, R), dtype=dtype, device=dev, generator=gen) * 0.1 * J + 0.03 * J
    peaks = (torch.exp(-0.5 * ((t - centre) / width) ** 2) + 0.02 * torch.randn((G, J, R), dtype=dtype, device=dev, generator=gen)).reshape(G * J, R).contiguous()
    for name, V in (("noise", noise), ("peaks", peaks)):
        ws = _ops.Workspace(dev, 256, R, dtype, unimodal_shape=(G, R, J))
        ref = None
        for variant in (0, 9, 11, 14):
            lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, variant)
            times = []
            for rep in range(4):
                aux, dual = torch.empty_like(V), V.clone()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _ops.prox_unimodal(aux, dual, off, G, R, J, True, ws); e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            if ref is None:
                ref = aux.clone()
            same = bool(torch.equal(aux, ref))
            ms = float(np.median(times[1:]))
            rec = {"dtype": str(dtype), "input": name, "variant": variant, "ms": ms, "bit_identical_to_first_variant": same,
                   "algorithmic_gbs": 3 * V.numel() * V.element_size() / ms / 1e6}
            out["runs"].append(rec); print(rec)
lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, DEFAULT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_unimodal.json"), "w"), indent=1)
