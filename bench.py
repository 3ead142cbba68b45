"""Benchmark of the AO-ADMM hot path (BASELINE.json metric: outer iterations/s at 16k slices + X-stream GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c3|c3f32|c4|c4w]

A *step* is one outer AO-ADMM iteration (B-, C-, A-mode ADMM updates + feasibility gaps + fit/loss with the small
device->host read the stopping rule needs; reference decomposition.py:945-1053) over the whole synthetic data set.
Workload at N=1: BASELINE.json config[2] (the one the metric is quoted on): non-negative PARAFAC2 + L1 on C,
16,384 ragged slices X_i (J_i x 1024, J_i ~ U[256, 2048]), rank 20, fp64 — ~155 GB of X resident in HBM; it is far
larger than the 126 MB L2, so no flush is needed between iterations.  With --gpus N the same 16,384 slices are
sharded over N ranks (strong scaling; one process per GPU under torchrun, NCCL).

Prints ONE JSON line (rank 0).  `value` = whole-job outer iterations/s with X resident in HBM; `roofline` = the
dominant X-stream kernel, timed live with CUDA events inside the timed steps; `cpu_baseline` = the oracle port
(NumPy restatement of the reference, all host threads) on a bounded slice sample; `e2e` = the same metric through the
public `cmf_aoadmm` call with HOST buffers (pack + H2D + fit + D2H inside the timed region).
`--impl reference` times the reference algorithm's CPU port (oracle/aoadmm_oracle.py) on the same workload's sample.
`cpu_baseline_torch` (both arms, N=1) = the same sample under the reference's PyTorch-backend flavour on the host
cores (oracle/aoadmm_torch_cpu.py, fp64 and fp32) — the north_star's "torch-CPU" column; absent-with-reason for
penalties that backend cannot run (Unimodality).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (slices I, K columns, (J_lo, J_hi), rank, dtype, penalties-kwargs, description)
    "c1": dict(I=4096, K=512, J=(256, 256), R=16, dtype="f64", kw=dict(non_negative=True),
               desc="BASELINE config[1]: nonneg CMF, 4096 slices 256x512, R=16, fp64"),
    "c2": dict(I=16384, K=1024, J=(256, 2048), R=20, dtype="f64",
               kw=dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}),
               desc="BASELINE config[2]: nonneg PARAFAC2 + L1 on C, 16384 ragged slices J_i x 1024 "
                    "(J_i~U[256,2048]), R=20, fp64"),
    "c3": dict(I=8192, K=256, J=(1024, 1024), R=8, dtype="f64",
               kw=dict(non_negative=True, parafac2=True, unimodal={1: True}, l2_norm_bound=[0, 1, 1]),
               desc="BASELINE config[3]: unimodal + L2Ball PARAFAC2, 8192 slices 1024x256, R=8, fp64"),
    "c3f32": dict(I=8192, K=256, J=(1024, 1024), R=8, dtype="f32",
                  kw=dict(non_negative=True, parafac2=True, unimodal={1: True}, l2_norm_bound=[0, 1, 1]),
                  desc="BASELINE config[3]: unimodal + L2Ball PARAFAC2, 8192 slices 1024x256, R=8, fp32"),
    "c4": dict(I=8192, K=2048, J=(512, 512), R=32, dtype="f64", kw=dict(non_negative=True),
               desc="BASELINE config[4] per-GPU share: nonneg CMF, 8192 slices 512x2048, R=32, fp64"),
}
# config[4] as the weak-scaling sweep BASELINE.json describes: 8192 slices PER GPU (65,536 at 8 GPUs: 550 GB of X, which
# no single GPU holds).  `value` stays outer iterations/s of the N-times larger problem; `slice_iterations_per_s` =
# value x slices is the aggregate that should grow with N.
# diagnostic (not a BASELINE config): config 1 at half the rank — the regime where the contraction is HBM-bound, not
# fp64-pipe-bound, i.e. where reading X once instead of twice pays in full
CONFIGS["c1r8"] = dict(CONFIGS["c1"], R=8, desc="config[1] shapes at R=8 (diagnostic): nonneg CMF, 4096 slices 256x512, fp64")
CONFIGS["c4w"] = dict(CONFIGS["c4"], weak=True,
                      desc="BASELINE config[4] weak scaling: nonneg CMF, 8192 slices 512x2048 PER GPU, R=32, fp64")


def slice_sizes(cfg, seed=1):
    lo, hi = cfg["J"]
    rs = np.random.RandomState(seed)
    return rs.randint(lo, hi + 1, size=cfg["I"]).astype(np.int64) if hi > lo else np.full(cfg["I"], lo, np.int64)


def shard_bounds(sizes, world):
    """Contiguous slice ranges balanced by row count."""
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world)))
    cuts.append(len(sizes))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def truth_factors(cfg, seed=1):
    rs = np.random.RandomState(seed + 1000)
    return rs.uniform(0.1, 1.1, size=(cfg["I"], cfg["R"])), rs.uniform(size=(cfg["K"], cfg["R"]))


def gen_device_data(cfg, sizes, lo, hi, dtype, device, seed=1):
    """X_i = (B_i o a_i) C^T + 20% noise for slices lo..hi, generated chunk-wise directly in HBM."""
    import torch

    from matcouply_b200 import _ops
    from matcouply_b200._engine import PackedMatrices

    K, R = cfg["K"], cfg["R"]
    A, C = truth_factors(cfg, seed)
    Ad = torch.as_tensor(A[lo:hi], dtype=dtype, device=device)
    Ct = torch.as_tensor(C, dtype=dtype, device=device).T.contiguous()
    loc = sizes[lo:hi]
    off = np.concatenate([[0], np.cumsum(loc)]).astype(np.int64)
    N = int(off[-1])
    ld = _ops.padded_ld(K, dtype)
    X = torch.empty((N, ld), dtype=dtype, device=device)
    if ld != K:
        X[:, K:] = 0
    gen = torch.Generator(device=device).manual_seed(seed * 7919 + lo)
    gor = torch.repeat_interleave(torch.arange(hi - lo, device=device), torch.as_tensor(loc, device=device))
    chunk = 1 << 16
    signal = float(np.sqrt(R / 3.0) * 0.35)  # ~ RMS of the noiseless entries; noise is 20 % of it
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        Bt = torch.rand((r1 - r0, R), dtype=dtype, device=device, generator=gen)
        Bt *= Ad[gor[r0:r1]]
        blk = Bt @ Ct
        blk += 0.2 * signal * torch.randn(blk.shape, dtype=dtype, device=device, generator=gen)
        X[r0:r1, :K] = blk
    return PackedMatrices(X, off, K)


def gen_host_sample(cfg, sizes, n_slices, seed=1):
    """The first n_slices of the same workload family as host NumPy arrays (CPU baseline / e2e input)."""
    A, C = truth_factors(cfg, seed)
    rs = np.random.RandomState(seed * 31 + 5)
    out = []
    signal = float(np.sqrt(cfg["R"] / 3.0) * 0.35)
    for i in range(n_slices):
        B = rs.uniform(size=(int(sizes[i]), cfg["R"]))
        M = (B * A[i]) @ C.T
        out.append(M + 0.2 * signal * rs.standard_normal(size=M.shape))
    return out


def make_regs(kw):
    from matcouply_b200.decomposition import _parse_all_penalties

    return _parse_all_penalties(
        non_negative=kw.get("non_negative"), lower_bound=None, upper_bound=None,
        l2_norm_bound=kw.get("l2_norm_bound"), unimodal=kw.get("unimodal"), parafac2=kw.get("parafac2"),
        l1_penalty=kw.get("l1_penalty"), tv_penalty=None, generalized_l2_penalty=None, svd="truncated_svd", regs=None,
        dual_init="random_uniform", aux_init="random_uniform", verbose=False)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def n_in(self, t0, t1):
        return sum(1 for t, _ in self.rows if t0 <= t <= t1)

    def stop(self, t0=None, t1=None, note=None):
        """Samples whose arrival time lies in [t0, t1] (the timed region); all samples if no window is given."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t0 is None or (t0 <= t <= t1 + 0.12))]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["window"] = note
        return out


def use_all_host_threads():
    """CPU legs: give BLAS / OpenMP / torch every core this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which would silently make the baseline single-threaded) and return the thread count ACTUALLY in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    used = 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=n)  # called as a function: stays in effect
        used = max([int(lib.get("num_threads", 1)) for lib in threadpool_info() if lib.get("user_api") == "blas"] or [1])
    except Exception:
        pass
    try:
        import torch

        torch.set_num_threads(n)
    except Exception:
        pass
    return used


def oracle_iter_seconds(mats, R, kw):
    """Reference algorithm on the CPU (oracle port): seconds per outer iteration = (t(3 its) - t(1 it)) / 2."""
    from oracle import aoadmm_oracle as O

    base = dict(random_state=0, tol=None, absolute_tol=None, **kw)
    O.ao_admm(mats[:2], R, n_iter_max=1, **base)  # warm-up (imports, BLAS threads)
    t0 = time.perf_counter()
    O.ao_admm(mats, R, n_iter_max=1, **base)
    t1 = time.perf_counter()
    O.ao_admm(mats, R, n_iter_max=3, **base)
    t2 = time.perf_counter()
    return max(((t2 - t1) - (t1 - t0)) / 2.0, 1e-9)


def reference_iter_seconds(mats, R, kw):
    """Seconds per outer iteration of the reference's OWN code on the host cores, and which code that was:
    ("reference") the unmodified `matcouply.decomposition.cmf_aoadmm` staged under oracle/_ref (oracle/stage_ref.py;
    NumPy backend through the tensorly stand-in, numba-jitted unimodal regression when numba is importable, as in the
    reference), or ("port") the oracle restatement when nothing is staged.  Same (t(3 its) - t(1 it)) / 2 rule."""
    ref = None
    if os.environ.get("B2_REFERENCE_ARM") != "port":  # "port" forces the restatement (A/B of the two CPU arms)
        try:
            from oracle.stage_ref import load_reference

            ref = load_reference()
        except Exception:
            ref = None
    if ref is None:
        return oracle_iter_seconds(mats, R, kw), "port"
    base = dict(random_state=0, tol=None, absolute_tol=None, return_errors=True, **kw)
    ref.cmf_aoadmm(mats[:2], R, n_iter_max=2, **base)  # warm-up (imports, BLAS threads, numba JIT)
    t0 = time.perf_counter()
    ref.cmf_aoadmm(mats, R, n_iter_max=1, **base)
    t1 = time.perf_counter()
    ref.cmf_aoadmm(mats, R, n_iter_max=3, **base)
    t2 = time.perf_counter()
    return max(((t2 - t1) - (t1 - t0)) / 2.0, 1e-9), "reference"


def torch_cpu_baseline(mats, R, kw, total_rows, n_slices_total):
    """The reference under TensorLy's PyTorch backend on the host cores (oracle/aoadmm_torch_cpu.py; float64 and the
    backend's default float32), same sample and same (t(3 its) - t(1 it)) / 2 rule as the NumPy column.  Only for the
    penalties that backend can run (no Unimodality); never raises — the line says why it is absent."""
    try:
        import torch

        from oracle.aoadmm_torch_cpu import ao_admm_torch_cpu

        pen = {k: v for k, v in kw.items() if k in ("non_negative", "parafac2", "l1_penalty")}
        if len(pen) != len(kw):
            return {"unavailable": "the reference's torch backend cannot run " +
                                   ", ".join(sorted(set(kw) - set(pen))) + " (penalties.py:1008-1009)"}
        rows = sum(m.shape[0] for m in mats)
        out = {"unit": "iter/s", "cores": int(torch.get_num_threads()), "kind": "port",
               "sample": f"first {len(mats)} of {n_slices_total} slices ({rows} of {total_rows} rows), scaled "
                         "linearly in rows"}
        ao_admm_torch_cpu(mats[:2], R, n_iter_max=1, **pen)  # warm-up
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            t0 = time.perf_counter()
            ao_admm_torch_cpu(mats, R, n_iter_max=1, dtype=dt, **pen)
            t1 = time.perf_counter()
            ao_admm_torch_cpu(mats, R, n_iter_max=3, dtype=dt, **pen)
            t2 = time.perf_counter()
            t_s = max(((t2 - t1) - (t1 - t0)) / 2.0, 1e-9)
            out["value_" + name] = 1.0 / (t_s * total_rows / rows)
        out["value"] = out["value_f64"]
        return out
    except Exception as exc:  # the torch column is an extra; it must never cost the bench line
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


E2E_ITERS = 50  # outer iterations of the timed end-to-end call (the parity horizon of BASELINE.json's north_star)


E2E_BYTES_PER_RANK = 64 << 30   # page-locked host data of the e2e leg per rank ...
E2E_BYTES_TOTAL = 128 << 30      # ... and over all ranks of the box (pinning and generating it is untimed but not free)


def e2e_budget_bytes(world):
    """Bytes of X per rank for the e2e leg: 64 GB at 1-2 ranks, 128 GB / N beyond, never more than ~40 % of the host's
    available memory (B2_E2E_GB overrides the per-rank figure)."""
    if os.environ.get("B2_E2E_GB"):
        return int(float(os.environ["B2_E2E_GB"]) * (1 << 30))
    budget = min(E2E_BYTES_PER_RANK, E2E_BYTES_TOTAL // max(world, 1))
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(line.split()[1]) for line in f if line.startswith("MemAvailable"))
        budget = min(budget, int(0.4 * avail_kb * 1024 / max(world, 1)))
    except Exception:
        pass
    return max(budget, 1 << 30)


def e2e_sample_size(cfg, es, budget_bytes):
    """Slices of the e2e leg: about `budget_bytes` of X in page-locked host memory."""
    mean_j = sum(cfg["J"]) / 2
    return int(max(8, min(cfg["I"], budget_bytes // (mean_j * cfg["K"] * es))))


def gen_pinned_sample(cfg, sizes, lo, hi, dtype, device):
    """Slices lo..hi of the workload (same generator as the HBM-resident data) as NumPy views of ONE page-locked
    host buffer: the input of the e2e leg."""
    import torch

    packed = gen_device_data(cfg, sizes, lo, hi, dtype, device)
    K = cfg["K"]
    host = torch.empty((packed.N, K), dtype=dtype, pin_memory=True)
    host.copy_(packed.X[:, :K])
    torch.cuda.synchronize()
    off = packed.row_offsets
    del packed
    torch.cuda.empty_cache()
    arr = host.numpy()
    return [arr[a:b] for a, b in zip(off[:-1], off[1:])], int(off[-1])


def run_e2e(cfg, sizes, dtype, es, device, world, rank, group, regs):
    """e2e leg: the public `cmf_aoadmm` call on page-locked HOST arrays, whole call timed (initial state drawn from
    the RandomState stream, H2D of X, E2E_ITERS outer iterations with the per-iteration diagnostics D2H, D2H of the
    factors).  With
    N ranks every rank uploads and fits its row-balanced shard of an N-times larger sample (NCCL all-reduces inside
    the call); the time is the max over ranks."""
    import torch

    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.distributed import make_shard

    kw = dict(cfg["kw"], random_state=0, tol=None, absolute_tol=None)
    S2 = min(cfg["I"], e2e_sample_size(cfg, es, e2e_budget_bytes(world)) * world)
    sub = dict(cfg, I=S2)
    sizes2 = sizes[:S2]
    if world > 1:
        import torch.distributed as dist

        sh = make_shard([int(j) for j in sizes2], rank, world)
        kw.update(process_group=group, shard=sh, gather_factors=False)
        lo, hi = sh.lo, sh.hi
    else:
        lo, hi = 0, S2
    views, _rows_local = gen_pinned_sample(sub, sizes2, lo, hi, dtype, device)
    rows2 = int(sizes2.sum())
    cmf_aoadmm(views, cfg["R"], n_iter_max=1, **kw)  # warm-up of the call path (allocator, staging buffers)
    times = {}
    for k in (5, E2E_ITERS):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = cmf_aoadmm(views, cfg["R"], n_iter_max=k, return_errors=True, **kw)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[k] = float(t.item())
        del out
    del views
    k_e2e, t_call = E2E_ITERS, times[E2E_ITERS]
    total_rows = int(sizes.sum())
    n_state = 1 + sum(1 if r.__class__.__name__ == "Parafac2" else 2 for r in regs[1])  # B + aux/dual uploads
    h2d = rows2 * cfg["K"] * es + n_state * rows2 * cfg["R"] * 8
    d2h = rows2 * cfg["R"] * 8 + (k_e2e + 1) * 64 * 8
    t_iter = (times[E2E_ITERS] - times[5]) / (E2E_ITERS - 5)
    return {
        "value": 1.0 / (t_call / k_e2e * total_rows / rows2), "unit": "iter/s",
        "h2d_bytes_per_step": int(h2d / k_e2e), "d2h_bytes_per_step": int(d2h / k_e2e),
        "unscaled": {"call_seconds": t_call, "iterations": k_e2e, "rows": rows2, "rows_full_workload": total_rows,
                     "x_gbytes": rows2 * cfg["K"] * es / 1e9, "iter_per_s_on_sample": k_e2e / t_call,
                     "seconds_per_iteration": t_iter, "fixed_seconds": times[5] - 5 * t_iter},
        "note": f"cmf_aoadmm(list of page-locked host arrays, n_iter_max={k_e2e}, return_errors=True) on the first "
                f"{S2} slices ({rows2} rows, {rows2 * cfg['K'] * es / 1e9:.1f} GB of X) sharded over {world} GPU(s): whole "
                f"call timed, max over ranks ({t_call:.3f} s: initial state drawn from the RandomState stream, H2D of X, {k_e2e} "
                f"outer iterations with the per-iteration diagnostics D2H, D2H of the factors), iterations/s = {k_e2e} / "
                f"t_call, scaled linearly in rows to the full workload (the sample is the largest page-locked host "
                f"buffer a few-minute run can stage: 64 GB per rank, 128 GB per box). The same call with n_iter_max=5 takes {times[5]:.3f} s, i.e. {1000 * t_iter:.2f} ms "
                f"per iteration + {times[5] - 5 * t_iter:.3f} s of fixed cost (upload + init + download)"}


def cpu_sample_size(cfg):
    # ~13 ms per (1152 x 1024, R=20) slice-iteration on 8 cores -> a few seconds per iteration
    rows_budget = 300_000 if cfg["kw"].get("unimodal") is None else 40_000
    mean_j = sum(cfg["J"]) / 2
    return int(max(8, min(cfg["I"], rows_budget // mean_j)))


def run_reference(args, cfg, sizes):
    """--impl reference: the reference's own CPU implementation — the unmodified package staged under oracle/_ref by
    `__graft_entry__.build()` (kind "reference"; the oracle port only if nothing is staged) — on a bounded slice
    sample, extrapolated linearly in the row count (every reference loop is per slice)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    S = cpu_sample_size(cfg)
    mats = gen_host_sample(cfg, sizes, S)
    rows = sum(m.shape[0] for m in mats)
    secs, kind = [], "port"
    for _ in range(max(1, min(args.steps, 3))):
        t_s, kind = reference_iter_seconds(mats, cfg["R"], cfg["kw"])
        secs.append(t_s)
    t_iter_sample = float(np.median(secs))
    total_rows = int(sizes.sum())
    t_full = t_iter_sample * total_rows / rows
    value = 1.0 / t_full
    line = {
        "impl": "reference", "metric": "AO-ADMM outer iterations per second", "value": value, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * t_full,
        "higher_is_better": True, "scaling": "weak" if cfg.get("weak") else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["desc"], "sample": f"first {S} slices ({rows} rows), time scaled by rows"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": cores, "kind": kind,
                         "sample": f"first {S} of {cfg['I']} slices ({rows} of {total_rows} rows); "
                                   f"{t_iter_sample:.3f} s per outer iteration on the sample, scaled linearly in rows"
                                   + ("; unmodified reference package (oracle/_ref) on the NumPy backend"
                                      if kind == "reference" else "; oracle port (no staged reference)")},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cpu_baseline_torch": torch_cpu_baseline(mats, cfg["R"], cfg["kw"], total_rows, cfg["I"]),
    }
    if cfg.get("weak"):
        line["slice_iterations_per_s"] = value * int(cfg["I"])
    emit(line)


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner) write to fd 1 too,
    so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the real stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def emit(line):
    out = _claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("B2_BENCH_CONFIG", "c2"), choices=sorted(CONFIGS))
    ap.add_argument("--slices", type=int, default=0, help="override the slice count (debug only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / e2e legs (debug only)")
    ap.add_argument("--graph", default="auto", choices=["auto", "off"],
                    help="auto = the engine's launch policy, as in cmf_aoadmm: the steady-state outer iteration of a "
                         "single-GPU CMF-type problem is replayed as a CUDA graph (same kernels, no launch gaps); the "
                         "per-kernel-family times then come from the same number of eagerly issued steps right after "
                         "the timed region; off = always issue the launches one by one")
    ap.add_argument("--x1", default="auto", choices=["auto", "on", "off"],
                    help="single-read fused X-stream pass: auto = the engine's policy (ranks <= 8), on = wherever the "
                         "kernel applies, off = always the two-pass schedule (A/B)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.slices:
        cfg["I"] = args.slices
        cfg["desc"] += f" [REDUCED to {args.slices} slices]"
    if cfg.get("weak"):  # per-GPU work fixed: the problem grows with the number of ranks
        n_ranks = int(os.environ.get("WORLD_SIZE", "1"))
        cfg["I"] *= n_ranks
        cfg["desc"] += f" [{cfg['I']} slices on {n_ranks} GPU(s)]"
    sizes = slice_sizes(cfg)
    if args.impl == "reference":
        return run_reference(args, cfg, sizes)

    import torch

    from matcouply_b200 import _lib
    from matcouply_b200._engine import AOADMMEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    dtype = torch.float64 if cfg["dtype"] == "f64" else torch.float32
    es = 8 if dtype == torch.float64 else 4
    lo, hi = shard_bounds(sizes, world)[rank]

    # does the shard fit? (X + ~7.6 state arrays of N x R incl. W padding); otherwise shrink and SAY SO
    free_b, _total_b = torch.cuda.mem_get_info(device)
    need = lambda a, b: int(sizes[a:b].sum()) * (cfg["K"] * es + 7.6 * cfg["R"] * es) + (3 << 30)  # noqa: E731
    reduced = False
    while need(lo, hi) > free_b and hi - lo > 16:
        hi = lo + (hi - lo) * 15 // 16
        reduced = True
    regs = make_regs(cfg["kw"])
    while True:
        try:
            packed = gen_device_data(cfg, sizes, lo, hi, dtype, device)
            eng = AOADMMEngine(packed, cfg["R"], regs, group=group, fuse_x1={"auto": None, "on": True, "off": False}[args.x1])
            eng.load_state_device(seed=rank)
            eng.prepare()
            break
        except torch.cuda.OutOfMemoryError:
            packed = eng = None
            torch.cuda.empty_cache()
            hi = lo + (hi - lo) * 7 // 8
            reduced = True
            if hi - lo < 16:
                raise
    x_bytes_local = packed.N * cfg["K"] * es

    use_graph = args.graph == "auto" and eng.graph_auto()
    graph_state = {"n": 0}

    def step():
        if use_graph and graph_state["n"] >= 2:  # steady state: replay the captured launch sequence
            return eng._read_diagnostics(eng.graph_iteration(True))
        graph_state["n"] += 1
        eng.outer_iteration()
        return eng.diagnostics()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs a few 100 ms to come up: start it before the warm-up, window the samples later
    for _ in range(max(args.warmup, 3) if use_graph else args.warmup):  # the third step captures the graph
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    eng.xstream_events = None if use_graph else {}
    launches0 = int(_lib.load().b2_launch_count())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_host1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = int(_lib.load().b2_launch_count()) - launches0
    if use_graph:
        # the replayed kernels cannot be bracketed with events: issue the same steps eagerly once more (untimed) for the
        # per-kernel-family times, and count the kernels of one step
        use_graph = False
        eng.xstream_events = {}
        launches0 = int(_lib.load().b2_launch_count())
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        launches = int(_lib.load().b2_launch_count()) - launches0
        use_graph = True
    ev = eng.xstream_events
    eng.xstream_events = None
    fam = {k: (float(np.mean([a.elapsed_time(b) for a, b in v])), len(v) / args.steps) for k, v in ev.items() if v}
    keys = ["y", "z", "fused", "rowpass", "polar", "unimodal", "local"]
    x_passes = 1 if eng.fused_x1 else 2
    tmax = torch.tensor([ms_total] + [fam.get(k, (0.0, 0.0))[0] for k in keys], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.barrier()
    vals = [float(v) for v in tmax.cpu()]
    ms_total = vals[0]
    fam = {k: (vals[1 + i], fam[k][1]) for i, k in enumerate(keys) if k in fam}
    t_y, t_z = fam.get("y", (None,))[0], fam.get("z", (None,))[0]
    ms_step = ms_total / args.steps
    window_note = "timed region"
    need = torch.tensor([1.0 if sampler.n_in(t_host0, t_host1 + 0.12) < 2 else 0.0], device=device)
    if world > 1:
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
    if float(need.item()) > 0:
        # the timed region is shorter than two nvidia-smi periods (100 ms): keep the SAME load running, untimed, for
        # ~0.7 s (the same number of steps on every rank: the steps contain collectives) and sample that instead
        n_extra = int(max(2, min(2000, 700.0 / max(ms_step, 1e-3))))
        t_host0 = time.perf_counter()
        for _ in range(n_extra):
            step()
        torch.cuda.synchronize()
        t_host1 = time.perf_counter()
        window_note = (f"{n_extra} more untimed steps of the same load right after the timed region "
                       "(region shorter than two 100 ms sampling periods)")
    clocks = sampler.stop(t_host0, t_host1, window_note)

    rows_rank0 = int(packed.N)
    del eng, packed
    torch.cuda.empty_cache()
    e2e = None
    if not args.no_cpu:
        e2e = run_e2e(cfg, sizes, dtype, es, device, world, rank, group, regs)  # collective: every rank takes part
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)" if "hbm_gbs" in peaks else "fallback 6650"
    # ---- roofline of the DOMINANT kernel of this config's step (largest time per step over the kernel families) ----
    # algorithmic bytes per launch: X-stream = the X shard once; the B-state kernels = the N x R arrays
    # they must read and write once (row pass: Y, V, T in / V', T' out in a middle pass; unimodal: V in, aux + dual out;
    # fused local loop: Y + x + (aux, dual) per penalty in, x + W + (aux, dual) out)
    nr_bytes = rows_rank0 * cfg["R"] * es
    n_b_pen = len(regs[1])
    # single-read fused pass: X once (its second read of every 64-row chunk is served by L2; the B-state and G_i = X_i^T B_i
    # traffic is not counted — conservative, like the side operands of the two-pass kernels)
    algo = {"y": x_bytes_local, "z": x_bytes_local, "fused": x_bytes_local,
            "rowpass": (3 + 2 * max(n_b_pen - 1, 0)) * nr_bytes,
            "unimodal": 3 * nr_bytes, "local": (4 + 4 * n_b_pen) * nr_bytes, "polar": 3 * cfg["R"] ** 2 * 8 * (hi - lo)}
    names = {"y": "xstream_y", "z": "xstream_z", "fused": "xfused_local_kernel", "rowpass": "pf2_rowpass_mma_kernel",
             "polar": "pf2_polar_reg_kernel",
             "unimodal": "unimodal_kernel", "local": "admm_local_mma_kernel"}
    kernels = {names[k]: {"ms_per_launch": t, "launches_per_step": n, "ms_per_step": t * n,
                          "algorithmic_bytes_per_launch": int(algo[k]), "gbs": algo[k] / t / 1e6 if t > 0 else None}
               for k, (t, n) in fam.items()}
    dom_key = max(fam, key=lambda k: fam[k][0] * fam[k][1])
    dom, t_dom = names[dom_key], fam[dom_key][0]
    achieved = algo[dom_key] / t_dom / 1e6  # GB/s
    traffic, traffic_src = None, None
    try:  # DRAM bytes per algorithmic byte from the committed `ncu --set full` capture, scaled to this launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # the capture at the FULL size of the headline config where there is one, else the 2 048-slice capture
        ent = tj.get(dom + "_fullsize") if (args.config == "c2" and not args.slices and dom + "_fullsize" in tj) else tj[dom]
        traffic = ent["ratio"] * algo[dom_key]
        traffic_src = (f"(dram__bytes_read.sum + dram__bytes_write.sum) / algorithmic bytes = {ent['ratio']:.4f} in "
                       f"the {ent.get('source', tj.get('source'))}, times this launch's algorithmic bytes")
    except Exception:
        pass
    # second roof of the fp64 X-stream kernels: the fp64 tensor pipe (DMMA.8x8x4), measured live (burst, idle GPU)
    fp64 = None
    if cfg["dtype"] == "f64":
        try:
            from matcouply_b200 import _ops

            flops, ms = _ops.microbench_flops(1, 4096)
            peak_tf = flops / ms / 1e9
            r_pad = (cfg["R"] + 7) // 8 * 8
            if "fused" in fam:  # both contractions (Y = X C and G = X^T B) in the one launch
                t_x, n_contr, kname = fam["fused"][0], 2, "xfused_local_kernel"
            else:
                t_x, n_contr, kname = max(t_y, t_z), 1, "xstream_y" if t_y >= t_z else "xstream_z"
            ach_tf = n_contr * 2.0 * r_pad * cfg["K"] * rows_rank0 / t_x / 1e9
            fp64 = {"kernel": kname, "achieved": ach_tf, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                    "note": f"{n_contr}*2*{r_pad}*K*N padded DMMA.8x8x4 flops per launch / CUDA-event time; peak = "
                            "b2_microbench_flops(DMMA) in this run at the clock of an otherwise idle GPU "
                            f"(the timed steps ran at {clocks.get('sm_mhz')} of {clocks.get('sm_max_mhz')} MHz)"}
        except Exception as exc:
            fp64 = {"unavailable": f"{type(exc).__name__}: {exc}"}
    line = {
        "metric": "AO-ADMM outer iterations per second", "value": 1000.0 / ms_step, "unit": "iter/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak" if cfg.get("weak") else "strong", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": cfg["desc"] + (" [REDUCED: shard did not fit HBM]" if reduced else ""),
                   "slices_total": int(cfg["I"]) if not reduced else int(hi - lo), "rows_rank0": rows_rank0,
                   "x_bytes_rank0": int(x_bytes_local), "x_passes_per_iteration": x_passes,
                   "l2_flush": "not needed: X shard >> 126 MB L2", "parallelism": f"slices sharded over {world} GPU(s)",
                   "launch_mode": ("CUDA-graph replay of the steady-state outer iteration (engine policy for single-GPU "
                                   "CMF-type problems; gpu_launches = kernels executed per timed step x steps, counted "
                                   "on eagerly issued steps)") if use_graph else "one launch per kernel"},
        "roofline": {"bound": "hbm+fp64" if (fp64 and "frac" in fp64 and dom_key in ("y", "z", "fused")) else "hbm",
                     "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "fp64": fp64, "kernels": kernels,
                     "xstream_y_ms": t_y, "xstream_z_ms": t_z,
                     "xstream_y_gbs": x_bytes_local / t_y / 1e6 if t_y else None,
                     "xstream_z_gbs": x_bytes_local / t_z / 1e6 if t_z else None,
                     "iteration_stream_gbs": x_passes * x_bytes_local / ms_step / 1e6},
        "gpu_launches": launches, "clocks": clocks,
    }
    if not args.no_cpu:
        # ---- cpu_baseline: oracle port on a bounded sample of the same workload (rank 0, N=1 only) ----
        S = cpu_sample_size(cfg)
        mats = gen_host_sample(cfg, sizes, S)
        rows = sum(m.shape[0] for m in mats)
        total_rows = int(sizes.sum())
        if world == 1:
            cores = use_all_host_threads()
            t_s, kind = reference_iter_seconds(mats, cfg["R"], cfg["kw"])
            line["cpu_baseline"] = {
                "value": 1.0 / (t_s * total_rows / rows), "unit": "iter/s", "cores": cores, "kind": kind,
                "sample": f"first {S} of {cfg['I']} slices ({rows} of {total_rows} rows): {t_s:.3f} s per outer "
                          f"iteration, scaled linearly in rows (every reference loop is per slice)"}
            line["cpu_baseline_torch"] = torch_cpu_baseline(mats, cfg["R"], cfg["kw"], total_rows, cfg["I"])
    if e2e is not None:
        line["e2e"] = e2e
    if cfg.get("weak"):
        line["slice_iterations_per_s"] = line["value"] * int(cfg["I"])
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
