"""Trajectory parity at the REAL widths of BASELINE.json's configs 1-4: the exact K / J-range / R / penalties of every
config, only the number of slices reduced (I = 12..32) so that the CPU oracle finishes in seconds (SURVEY.md §8d:
"scaled-down twins, same K_ref/J/R/penalties").  These are the kernel instantiations bench.py runs
(`xstream_y_kernel<double,1,1,3,0>`, `pf2_polar_reg_kernel<double,20>`, `pf2_rowpass_mma_kernel<double,2,1,...>`,
`admm_local_mma_kernel` at R = 16 / 32, the unimodal kernel at J = 1024 ...), compared over 50 outer iterations with the
oracle run live on the same seeded inputs (oracle/aoadmm_oracle.py; pinned to the unmodified reference, bit-identical to
/root/reference/src/matcouply/decomposition.py:662 at K = 1024, R = 20).

Tolerances (BASELINE.json north_star): factor matrices within 1e-8 relative at every checkpoint of the first 50
iterations in fp64, losses within 1e-6 (asserted at 1e-8), same stopping iteration and message.  fp32 (config 3): the
fp32 kernels on fp32 inputs against the fp64 oracle on the fp32-rounded inputs, at single-precision tolerance.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHECKPOINTS = (1, 2, 3, 5, 10, 20, 35, 50)
# slices of the twin per config: enough rows that every CTA-per-slice kernel sees several slices, small enough for
# the oracle (c2: ~0.3 s per outer iteration on 8 host threads)
TWIN_SLICES = {"c1": 32, "c2": 16, "c3": 16, "c4": 12}


class _Checkpoints(list):
    """`trajectory` sink of the oracle that keeps the factors of the checkpoint iterations only."""

    def __init__(self, keep):
        super().__init__()
        self.keep, self.n = set(keep), 0

    def append(self, snap):
        self.n += 1
        if self.n in self.keep:
            super().append((self.n, snap["A"], np.concatenate(snap["B_is"], 0), snap["C"]))


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def twin(name, dtype=np.float64):
    import bench

    cfg = dict(bench.CONFIGS[name])
    cfg["I"] = TWIN_SLICES[name]
    sizes = bench.slice_sizes(cfg)
    X = bench.gen_host_sample(cfg, sizes, cfg["I"])
    if dtype != np.float64:
        X = [x.astype(dtype) for x in X]
    return cfg, X


_ORACLE_CACHE = {}


def oracle_trajectory(name, X, cfg):
    """(checkpoints, losses) of the 50-iteration oracle run; cached per config (the multi-rank tests reuse it)."""
    from oracle import aoadmm_oracle as O

    if name not in _ORACLE_CACHE:
        traj = _Checkpoints(CHECKPOINTS)
        o = O.ao_admm([np.asarray(x, dtype=np.float64) for x in X], cfg["R"], n_iter_max=max(CHECKPOINTS), tol=None,
                      absolute_tol=None, random_state=0, trajectory=traj, **cfg["kw"])
        _ORACLE_CACHE[name] = (list(traj), np.asarray(o["regularized_loss"]), np.asarray(o["rec_errors"]))
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4"])
def test_fp64_trajectory_at_baseline_width(name):
    from matcouply_b200 import cmf_aoadmm

    cfg, X = twin(name)
    traj, losses, recs = oracle_trajectory(name, X, cfg)
    worst = 0.0
    for k, A_o, B_o, C_o in traj:
        cmf, diag = cmf_aoadmm(X, cfg["R"], n_iter_max=k, tol=None, absolute_tol=None, random_state=0,
                               return_errors=True, **cfg["kw"])
        _, (A, B_is, C) = cmf
        errs = (rel(A, A_o), rel(np.concatenate(B_is, 0), B_o), rel(C, C_o))
        worst = max(worst, *errs)
        assert max(errs) < 1e-8, (name, k, errs)
    # the last run is the 50-iteration one: its per-iteration losses against the oracle's
    np.testing.assert_allclose(diag.regularized_loss, losses, rtol=1e-8)
    np.testing.assert_allclose(diag.rec_errors, recs, rtol=1e-8)
    print(f"[width parity] {name} fp64: K={cfg['K']} R={cfg['R']} I={cfg['I']} rows={sum(x.shape[0] for x in X)}: "
          f"worst relative factor difference over checkpoints {CHECKPOINTS} = {worst:.3e}, "
          f"final loss rel diff = {abs(diag.regularized_loss[-1] - losses[-1]) / losses[-1]:.3e}")


@pytest.mark.parametrize("rank", [16, 8])
def test_fp64_trajectory_config1_single_read_pass(rank, monkeypatch):
    """SURVEY.md §8 row X1 at the config-1 width (K = 512, J = 256; R = 16 and the HBM-bound R = 8): the single-read
    fused pass (csrc/xfused.cu, forced on: at R = 16 the engine's own policy picks the two-pass schedule) against the
    oracle, 1e-8 per checkpoint over 50 iterations, and really ONE pass over X per outer iteration."""
    from matcouply_b200 import _engine, _ops, cmf_aoadmm
    from oracle import aoadmm_oracle as O

    cfg, X = twin("c1")
    traj = _Checkpoints(CHECKPOINTS)
    o = O.ao_admm(X, rank, n_iter_max=max(CHECKPOINTS), tol=None, absolute_tol=None, random_state=0, trajectory=traj,
                  **cfg["kw"])
    calls = {"fused": 0, "y": 0, "z": 0}
    for key in ("fused", "y", "z"):
        real = getattr(_ops, {"fused": "xstream_fused_local", "y": "xstream_y", "z": "xstream_z"}[key])
        monkeypatch.setattr(_ops, real.__name__, lambda *a, _k=key, _f=real, **kw: (
            calls.__setitem__(_k, calls[_k] + 1), _f(*a, **kw))[1])
    default = _engine.FUSION_DEFAULTS["x1"]
    _engine.FUSION_DEFAULTS.update(x1=True)
    try:
        worst = 0.0
        for k, A_o, B_o, C_o in traj:
            calls.update(fused=0, y=0, z=0)
            cmf, diag = cmf_aoadmm(X, rank, n_iter_max=k, tol=None, absolute_tol=None, random_state=0,
                                   return_errors=True, use_cuda_graph=False, **cfg["kw"])  # eager: the launches are counted
            assert (calls["fused"], calls["y"], calls["z"]) == (k, 1, 0)  # + the Y pass of the initial fit
            _, (A, B_is, C) = cmf
            errs = (rel(A, A_o), rel(np.concatenate(B_is, 0), B_o), rel(C, C_o))
            worst = max(worst, *errs)
            assert max(errs) < 1e-8, (rank, k, errs)
        # the same 50 iterations with the engine's own launch policy (CUDA-graph replay from the third iteration on):
        # same kernels on the same buffers, bit-identical to the eager run
        cmf_g, diag_g = cmf_aoadmm(X, rank, n_iter_max=max(CHECKPOINTS), tol=None, absolute_tol=None, random_state=0,
                                   return_errors=True, **cfg["kw"])
        np.testing.assert_array_equal(cmf_g[1][2], cmf[1][2])
        np.testing.assert_array_equal(np.concatenate(cmf_g[1][1], 0), np.concatenate(cmf[1][1], 0))
        np.testing.assert_array_equal(diag_g.regularized_loss, diag.regularized_loss)
    finally:
        _engine.FUSION_DEFAULTS.update(x1=default)
    np.testing.assert_allclose(diag.regularized_loss, o["regularized_loss"], rtol=1e-8)
    np.testing.assert_allclose(diag.rec_errors, o["rec_errors"], rtol=1e-8)
    print(f"[width parity] c1 single-read pass, R={rank}: worst relative factor difference over checkpoints "
          f"{CHECKPOINTS} = {worst:.3e}")


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4"])
def test_same_stopping_iteration_at_baseline_width(name):
    """Run to the stopping rule (loosened so that it fires within a few dozen iterations): same `n_iter`, message and
    flags as the oracle, final loss within 1e-6."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    cfg, X = twin(name)
    kw = dict(cfg["kw"], random_state=0, n_iter_max=30, tol=2e-3, feasibility_tol=0.3)
    o = O.ao_admm(X, cfg["R"], **kw)
    cmf, diag = cmf_aoadmm(X, cfg["R"], return_errors=True, **kw)
    assert diag.n_iter == o["n_iter"], (diag.n_iter, o["n_iter"])
    assert diag.message == o["message"]
    assert str(diag.satisfied_stopping_condition) == str(o["satisfied_stopping_condition"])
    assert abs(diag.regularized_loss[-1] - o["regularized_loss"][-1]) <= 1e-6 * abs(o["regularized_loss"][-1])
    print(f"[width parity] {name}: stopped after {diag.n_iter} iterations ({diag.message}) like the oracle")


def test_fp32_config3_at_baseline_width():
    """Config 3 in fp32 (BASELINE: "fp32 and fp64"): the fp32 kernels against the fp64 oracle on the fp32-rounded
    inputs.  Single-precision tolerance: the iterates of two precisions drift apart with the iteration count, so the
    bound is on the first 20 iterations' factors and on the 50-iteration loss."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    cfg, X32 = twin("c3", np.float32)
    X64 = [x.astype(np.float64) for x in X32]
    traj = _Checkpoints((5, 20, 50))
    o = O.ao_admm(X64, cfg["R"], n_iter_max=50, tol=None, absolute_tol=None, random_state=0, trajectory=traj,
                  **cfg["kw"])
    report = []
    for k, A_o, B_o, C_o in traj:
        cmf, diag = cmf_aoadmm(X32, cfg["R"], n_iter_max=k, tol=None, absolute_tol=None, random_state=0,
                               return_errors=True, **cfg["kw"])
        _, (A, B_is, C) = cmf
        errs = (rel(A, A_o), rel(np.concatenate(B_is, 0), B_o), rel(C, C_o))
        report.append((k, errs))
        if k <= 20:
            assert max(errs) < 2e-2, (k, errs)
    loss_err = abs(diag.regularized_loss[-1] - o["regularized_loss"][-1]) / o["regularized_loss"][-1]
    print(f"[width parity] c3 fp32 vs fp64 oracle on fp32-rounded inputs: {report}, final loss rel diff {loss_err:.3e}")
    assert loss_err < 1e-3, loss_err
    assert np.all(np.isfinite(diag.regularized_loss))
