"""A/B of the PARAFAC2 row-pass kernels inside one process: B2_OPT_PF2_ROWPASS_MMA = 1 (general DMMA tile kernel) vs 2
(steady-state specialisation, csrc/pf2_rowpass_v2.cuh) on a config-2 shard; CUDA-event time per launch, alternating.
    python tools/ab_rowpass.py [slices]  ->  gpurun_out/ab_rowpass.json"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from matcouply_b200 import _lib  # noqa: E402
from matcouply_b200._engine import AOADMMEngine  # noqa: E402

slices = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = dict(bench.CONFIGS["c2"]); cfg["I"] = slices
sizes = bench.slice_sizes(cfg)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
packed = bench.gen_device_data(cfg, sizes, 0, slices, torch.float64, dev)
eng = AOADMMEngine(packed, cfg["R"], bench.make_regs(cfg["kw"]))
eng.load_state_device(seed=0); eng.prepare()
lib = _lib.load()
for _ in range(2):
    eng.outer_iteration(); eng.diagnostics()
out = {"slices": slices, "rows": int(packed.N), "nr_bytes": int(packed.N) * cfg["R"] * 8, "runs": []}
for rep in range(3):
    for opt in (1, 2):
        lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, opt)
        eng.xstream_events = {}
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.outer_iteration(); eng.diagnostics()
        e1.record(); torch.cuda.synchronize()
        ev = eng.xstream_events; eng.xstream_events = None
        t = [a.elapsed_time(b) for a, b in ev["rowpass"]]
        per_pos = [float(np.mean(t[i::5])) for i in range(5)]  # first, 3 middle, last pass of a B-update
        out["runs"].append({"option": opt, "ms_per_iteration": e0.elapsed_time(e1) / 3, "rowpass_ms_by_pass": per_pos,
                            "rowpass_ms_mean": float(np.mean(t)), "middle_pass_gbs": 5 * out["nr_bytes"] / np.mean(per_pos[1:4]) / 1e6})
        print(out["runs"][-1])
lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, _lib.PF2_ROWPASS_DEFAULT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_rowpass.json"), "w"), indent=1)
