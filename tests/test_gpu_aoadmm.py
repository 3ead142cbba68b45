"""End-to-end parity of ``matcouply_b200.cmf_aoadmm`` (CUDA path, called like the reference) against

* golden trajectories produced by the UNMODIFIED reference (tests/golden/traj_*.npz, recipe: oracle/gen_golden.py),
* the oracle run live on fresh seeded inputs.

Tolerances (BASELINE.json north_star): factor matrices within 1e-8 relative per iteration for the first 50
iterations in fp64, final relative loss within 1e-6, same stopping iteration.
"""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
# the init_* cases start from a host LAPACK SVD whose signs may depend on the CPU the test runs on: they are compared
# with the oracle run live on the same box (test_svd_inits_match_live_oracle) instead of with the stored trajectory
CASES = sorted(n for n in (os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(HERE, "golden", "traj_*.npz")))
               if not n.startswith("init_"))


def load_case(name):
    g = np.load(os.path.join(HERE, "golden", f"traj_{name}.npz"), allow_pickle=False)
    off = g["row_offsets"]
    X = [g["X"][a:b] for a, b in zip(off[:-1], off[1:])]
    kw = json.loads(str(g["kwargs"]))
    for key in ("l1_penalty", "non_negative", "unimodal", "l2_norm_bound", "lower_bound", "upper_bound", "tv_penalty"):
        if isinstance(kw.get(key), dict):
            kw[key] = {int(k): v for k, v in kw[key].items()}
    if isinstance(kw.get("generalized_l2_penalty"), dict):  # norm matrices are stored as nested lists
        kw["generalized_l2_penalty"] = {int(k): np.asarray(v, dtype=np.float64)
                                        for k, v in kw["generalized_l2_penalty"].items()}
    return g, X, int(g["rank"]), kw


def product_kwargs(kw):
    """`regs_spec` (JSON-safe description of an explicit `regs` argument) -> matcouply_b200 penalty objects."""
    from matcouply_b200 import penalties as P
    from oracle.aoadmm_oracle import regs_from_spec

    kw = dict(kw)
    if "regs_spec" in kw:
        kw["regs"] = regs_from_spec(kw.pop("regs_spec"), P)
    return kw


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", CASES)
def test_trajectory_matches_reference(name):
    """Per-iteration factors for the first 50 outer iterations: chain 1-iteration... no: run k iterations from the
    same seed (the run is deterministic) at a few checkpoints and compare with the reference trajectory."""
    from matcouply_b200 import cmf_aoadmm

    g, X, rank, kw = load_case(name)
    n_traj = g["A_traj"].shape[0] if g["A_traj"].ndim == 3 else 0
    if n_traj == 0:
        pytest.skip("no trajectory stored")
    kw = product_kwargs(kw)
    kw.pop("n_iter_max", None)
    worst = 0.0
    for k in sorted({1, 2, 3, 5, 10, 20, 35, min(50, n_traj)}):
        if k > n_traj:
            continue
        cmf = cmf_aoadmm(X, rank, n_iter_max=k, tol=None, absolute_tol=None, **kw)
        _, (A, B_is, C) = cmf
        errs = (rel(A, g["A_traj"][k - 1]), rel(np.concatenate(B_is, 0), g["B_traj"][k - 1]),
                rel(C, g["C_traj"][k - 1]))
        worst = max(worst, *errs)
        assert max(errs) < 1e-8, (name, k, errs)
    print(f"{name}: worst relative factor difference over checkpoints = {worst:.3e}")


@pytest.mark.parametrize("name", CASES)
def test_full_run_matches_reference(name):
    """Same stopping iteration and message, final loss within 1e-6 relative, diagnostics lists element-wise close."""
    from matcouply_b200 import cmf_aoadmm

    g, X, rank, kw = load_case(name)
    kw = product_kwargs(kw)
    cmf, admm, diag = cmf_aoadmm(X, rank, return_errors=True, return_admm_vars=True, **kw)
    assert diag.n_iter == int(g["n_iter"]), (diag.n_iter, int(g["n_iter"]))
    assert diag.message == str(g["message"])
    assert str(diag.satisfied_stopping_condition) == str(g["satisfied_stopping_condition"])
    assert str(diag.satisfied_feasibility_condition) == str(g["satisfied_feasibility_condition"])
    ref_loss = g["regularized_loss"]
    assert len(diag.regularized_loss) == len(ref_loss)
    assert abs(diag.regularized_loss[-1] - ref_loss[-1]) <= 1e-6 * abs(ref_loss[-1])
    np.testing.assert_allclose(diag.regularized_loss[:51], ref_loss[:51], rtol=1e-8)
    np.testing.assert_allclose(diag.rec_errors[:51], g["rec_errors"][:51], rtol=1e-8)
    gaps = np.array([[v for mode in it for v in mode] for it in diag.feasibility_gaps], dtype=np.float64)
    assert gaps.shape == g["feasibility_gaps"].shape
    if gaps.size:
        np.testing.assert_allclose(gaps[:51], g["feasibility_gaps"][:51], rtol=1e-6, atol=1e-12)
    _, (A, B_is, C) = cmf
    if diag.n_iter <= 130:  # short runs stay within round-off of the reference up to the end
        assert rel(A, g["A"]) < 1e-7 and rel(C, g["C"]) < 1e-7 and rel(np.concatenate(B_is, 0), g["B"]) < 1e-7
        # ADMM variables in the reference's ADMMVars layout
        for m in range(3):
            for n, (aux, dual) in enumerate(zip(admm.auxes[m], admm.duals[m])):
                key = f"m{m}_r{n}"
                d = np.concatenate(dual, 0) if m == 1 else dual
                assert rel(d, g["dual_" + key]) < 1e-6, key
                if isinstance(aux, tuple):
                    assert rel(aux[1], g["aux_" + key + "_coord"]) < 1e-6
                    assert rel(np.concatenate(aux[0], 0), g["aux_" + key + "_basis"]) < 1e-6
                else:
                    a = np.concatenate(aux, 0) if m == 1 else aux
                    assert rel(a, g["aux_" + key]) < 1e-6, key


def test_return_errors_false_skips_loss_like_reference():
    """Stopping iteration with return_errors=False (loss skipped on infeasible iterations) equals the oracle's."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(5)
    X = [rs.uniform(size=(J, 9)) for J in (7, 9, 8, 12, 6)]
    kw = dict(non_negative=True, parafac2=True, random_state=3, n_iter_max=400, tol=1e-6)
    o = O.ao_admm(X, 2, return_errors=False, **kw)
    cmf, diag = cmf_aoadmm(X, 2, return_errors=False, return_admm_vars=False, **kw), None
    o2 = O.ao_admm(X, 2, return_errors=True, **kw)
    cmf2, diag2 = cmf_aoadmm(X, 2, return_errors=True, **kw)
    assert diag2.n_iter == o2["n_iter"]
    assert rel(cmf2[1][0], o2["A"]) < 1e-6
    assert rel(cmf[1][0], o["A"]) < 1e-6


def test_live_oracle_ragged_fp64_and_fp32():
    """Fresh seeded ragged problem (scaled-down twin of BASELINE config 2): fp64 within 1e-8 after 30 iterations,
    fp32 within single-precision tolerance of the fp64 oracle on fp32-rounded inputs."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(42)
    I, K, R = 24, 40, 6
    Js = rs.randint(R, 70, size=I)
    A, C = rs.uniform(0.1, 1.1, size=(I, R)), rs.uniform(size=(K, R))
    X = []
    for i, J in enumerate(Js):
        M = (rs.uniform(size=(J, R)) * A[i]) @ C.T
        X.append(M + 0.1 * rs.standard_normal(size=M.shape))
    kw = dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}, random_state=0, n_iter_max=30, tol=None,
              absolute_tol=None)
    o = O.ao_admm(X, R, **kw)
    _, (Ag, Bg, Cg) = cmf_aoadmm(X, R, **kw)
    assert rel(Ag, o["A"]) < 1e-8 and rel(Cg, o["C"]) < 1e-8
    assert rel(np.concatenate(Bg, 0), np.concatenate(o["B_is"], 0)) < 1e-8
    X32 = [x.astype(np.float32) for x in X]
    o32 = O.ao_admm([x.astype(np.float64) for x in X32], R, **kw)
    _, (A32, B32, C32) = cmf_aoadmm(X32, R, **kw)
    assert rel(A32, o32["A"]) < 5e-3 and rel(C32, o32["C"]) < 5e-3


def test_edge_cases_and_api_errors():
    from matcouply_b200 import cmf_aoadmm, parafac2_aoadmm
    from matcouply_b200.penalties import NonNegativity
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(1)
    X = [rs.uniform(size=(J, 6)) for J in (5, 1, 9, 3)]  # includes a single-row slice
    cmf, diag = cmf_aoadmm(X, 2, n_iter_max=0, non_negative=True, return_errors=True, random_state=0)
    o = O.ao_admm(X, 2, n_iter_max=0, non_negative=True, random_state=0)
    assert diag.n_iter == 0 and len(diag.regularized_loss) == 1
    assert abs(diag.regularized_loss[0] - o["regularized_loss"][0]) < 1e-12
    # 3-D array input, explicit regs list, zero iterations produce the init
    T = rs.uniform(size=(4, 7, 5))
    cmf2 = cmf_aoadmm(T, 3, n_iter_max=3, regs=[[NonNegativity()], [], []], random_state=1)
    o2 = O.ao_admm(list(T), 3, n_iter_max=3, regs=[[O.NonNeg()], [], []], random_state=1, tol=None, absolute_tol=None)
    assert rel(cmf2[1][0], o2["A"]) < 1e-9
    with pytest.raises(TypeError):
        cmf_aoadmm(X, 2, regs=[[1], [], []])
    with pytest.raises(ValueError):
        cmf_aoadmm(X, 2, constant_feasibility_penalty="C")
    with pytest.raises(ValueError):
        cmf_aoadmm(X, 2, l1_penalty=[1, 2])
    with pytest.raises(ValueError):
        cmf_aoadmm(X, 2, init="nonsense")
    out = parafac2_aoadmm(X, 2, n_iter_max=2, non_negative=True, random_state=0, return_admm_vars=True)
    assert len(out) == 2 and isinstance(out[1].auxes[1][0], tuple)


@pytest.mark.parametrize("kw", [
    dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}),
    dict(non_negative=True, parafac2=True, unimodal={1: True}, l2_norm_bound=[0, 1, 1]),
    dict(non_negative=True, lower_bound={2: 0.0}, upper_bound={2: 0.9}),
])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_kernels_match_stepwise_kernels(kw, dtype):
    """The fused kernels (b2_admm_local, b2_pf2_rowpass) and the one-kernel-per-step path give the same iterates."""
    from matcouply_b200 import _engine, cmf_aoadmm

    rs = np.random.RandomState(17)
    I, K, R = 9, 33, 5
    Js = list(rs.randint(R, 150, size=I - 1)) + [70]
    X = [rs.uniform(size=(J, K)).astype(dtype) for J in Js]
    out = []
    for fused in (True, False):
        _engine.FUSION_DEFAULTS.update(local=fused, pf2=fused)
        try:
            cmf, admm = cmf_aoadmm(X, R, n_iter_max=6, tol=None, absolute_tol=None, random_state=2,
                                   return_admm_vars=True, **kw)
        finally:
            _engine.FUSION_DEFAULTS.update(local=True, pf2=True)
        out.append((cmf, admm))
    tol = 1e-9 if dtype == np.float64 else 2e-3
    (c1, a1), (c2, a2) = out
    assert rel(c1[1][0], c2[1][0]) < tol and rel(c1[1][2], c2[1][2]) < tol
    assert rel(np.concatenate(c1[1][1], 0), np.concatenate(c2[1][1], 0)) < tol
    for m in range(3):
        for x1, x2 in zip(a1.duals[m], a2.duals[m]):
            assert rel(np.concatenate(x1, 0) if m == 1 else x1, np.concatenate(x2, 0) if m == 1 else x2) < 10 * tol
        for x1, x2 in zip(a1.auxes[m], a2.auxes[m]):
            if isinstance(x1, tuple):
                assert rel(x1[1], x2[1]) < 10 * tol and rel(np.concatenate(x1[0], 0), np.concatenate(x2[0], 0)) < 10 * tol
            else:
                assert rel(np.concatenate(x1, 0) if m == 1 else x1, np.concatenate(x2, 0) if m == 1 else x2) < 10 * tol


@pytest.mark.parametrize("name", ["c0_readme", "c1_nn_cmf", "c2_nn_pf2_l1_ragged", "c3_unimodal_l2ball_pf2"])
def test_cuda_graph_replay_equals_eager_launches(name, monkeypatch):
    """The steady-state outer iteration replayed as a CUDA graph launches the same kernels on the same buffers as the
    eager loop: factors, ADMM variables and diagnostics must be bit-identical.  (A replayed PARAFAC2 iteration starts
    its first polar step cold every time, like the eager loop at its default B2_POLAR_COLD_EVERY=1.)"""
    from matcouply_b200 import cmf_aoadmm

    monkeypatch.setenv("B2_POLAR_COLD_EVERY", "1")

    g, X, rank, kw = load_case(name)
    kw = dict(kw)
    kw.pop("n_iter_max", None)
    outs = []
    for graph in (False, True):
        cmf, admm, diag = cmf_aoadmm(X, rank, n_iter_max=12, return_errors=True, return_admm_vars=True,
                                     use_cuda_graph=graph, **dict(kw, tol=None, absolute_tol=None))
        outs.append((cmf, admm, diag))
    (c0, a0, d0), (c1, a1, d1) = outs
    np.testing.assert_array_equal(c0[1][0], c1[1][0])
    np.testing.assert_array_equal(np.concatenate(c0[1][1], 0), np.concatenate(c1[1][1], 0))
    np.testing.assert_array_equal(c0[1][2], c1[1][2])
    np.testing.assert_array_equal(d0.regularized_loss, d1.regularized_loss)
    np.testing.assert_array_equal(d0.rec_errors, d1.rec_errors)
    for m in (0, 2):
        for x, y in zip(a0.duals[m], a1.duals[m]):
            np.testing.assert_array_equal(x, y)


def test_user_defined_python_penalty_is_bridged():
    """A penalty written against the reference's protocol in plain NumPy (examples/plot_custom_penalty.py:220-231
    style) runs inside cmf_aoadmm through the PEN_HOST bridge and reproduces the built-in CUDA penalty exactly."""
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200 import penalties as P

    class MyNonNeg(P.HardConstraintMixin, P.RowVectorPenalty):
        def factor_matrix_row_update(self, factor_matrix_row, feasibility_penalty, aux_row):
            return np.maximum(factor_matrix_row, 0)

    class MyL1(P.MatrixPenalty):
        def __init__(self, strength, **kw):
            super().__init__(**kw)
            self.strength = strength

        def factor_matrix_update(self, factor_matrix, feasibility_penalty, aux):
            return np.sign(factor_matrix) * np.maximum(np.abs(factor_matrix) - self.strength / feasibility_penalty, 0)

        def penalty(self, x):
            xs = [x] if isinstance(x, np.ndarray) else x
            return self.strength * sum(np.abs(xi).sum() for xi in xs)

    rs = np.random.RandomState(3)
    X = [rs.uniform(size=(J, 10)) for J in (6, 9, 7, 12)]
    common = dict(random_state=1, n_iter_max=25, tol=None, absolute_tol=None, return_errors=True)
    ref, dref = cmf_aoadmm(X, 3, regs=[[P.NonNegativity()], [P.L1Penalty(0.05)], [P.NonNegativity()]], **common)
    got, dgot = cmf_aoadmm(X, 3, regs=[[MyNonNeg()], [MyL1(0.05)], [MyNonNeg()]], **common)
    for a, b in zip((ref[1][0], np.concatenate(ref[1][1]), ref[1][2]), (got[1][0], np.concatenate(got[1][1]), got[1][2])):
        assert rel(b, a) < 1e-12
    np.testing.assert_allclose(dgot.regularized_loss, dref.regularized_loss, rtol=1e-12)


def test_config2_width_properties_on_device_resident_input():
    """BASELINE config 2 at full WIDTH (K = 1024, R = 20, ragged J_i in [256, 2048]) on 192 slices (1.8 GB) held as a
    device-resident PackedMatrices: too big for the oracle, so size-independent properties are checked — orthonormal
    PARAFAC2 bases, B_i = P_i Delta up to the reported feasibility gap, non-negative auxiliary variables, the expanded
    fit term against the naive residual, monotone loss once feasible, and bit-identical repeatability."""
    import bench
    from matcouply_b200 import cmf_aoadmm

    cfg = dict(bench.CONFIGS["c2"])
    cfg["I"] = 192
    sizes = bench.slice_sizes(cfg)
    dev_ = torch.device("cuda", 0)
    packed = bench.gen_device_data(cfg, sizes, 0, cfg["I"], torch.float64, dev_)
    kw = dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}, random_state=0, n_iter_max=12, tol=None,
              absolute_tol=None, return_errors=True, return_admm_vars=True)
    cmf, admm, diag = cmf_aoadmm(packed, cfg["R"], **kw)
    _, (A, B_is, C) = cmf
    assert np.all(np.isfinite(diag.regularized_loss)) and diag.n_iter == 12
    assert diag.regularized_loss[-1] < diag.regularized_loss[1]
    bases, delta = admm.auxes[1][0]
    gap_pf2 = diag.feasibility_gaps[-1][1][0]
    num = den = 0.0
    for i in (0, 1, 57, 191):
        P = bases[i]
        np.testing.assert_allclose(P.T @ P, np.eye(cfg["R"]), atol=1e-9)
        num += np.sum((P @ delta - B_is[i]) ** 2)
        den += np.sum(B_is[i] ** 2)
    assert np.sqrt(num / den) < 3 * gap_pf2 + 1e-12  # sampled slices vs the all-slice gap
    assert min(a.min() for a in admm.auxes[1][1]) >= 0 and admm.auxes[0][0].min() >= 0
    # naive residual on the device vs the expanded fit formula (decomposition.py:446-452)
    X, off = packed.X[:, :packed.K], packed.row_offsets
    Ct = torch.as_tensor(C, device=dev_).T.contiguous()
    sse = normx = 0.0
    for i in range(cfg["I"]):
        Xi = X[off[i]:off[i + 1]]
        Bi = torch.as_tensor(B_is[i] * A[i], device=dev_)
        sse += float(((Xi - Bi @ Ct) ** 2).sum())
        normx += float((Xi ** 2).sum())
    np.testing.assert_allclose(diag.rec_errors[-1], np.sqrt(sse / normx), rtol=1e-9)
    cmf2, _, diag2 = cmf_aoadmm(packed, cfg["R"], **kw)
    np.testing.assert_array_equal(cmf2[1][2], C)
    np.testing.assert_array_equal(np.asarray(diag2.regularized_loss), np.asarray(diag.regularized_loss))


def test_new_penalties_float32_inputs_track_float64():
    """float32 inputs run the fp32 kernels end to end (also for GeneralizedL2 / UnitSimplex / TV / Parafac2 options):
    the factors stay within single-precision distance of the float64 run after 15 iterations."""
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200 import penalties as P

    rs = np.random.RandomState(12)
    I, K, R, J = 7, 16, 3, 12
    A, C = rs.uniform(0.2, 1.2, size=(I, R)), rs.uniform(size=(K, R))
    X = [(rs.uniform(size=(J, R)) * a) @ C.T + 0.05 * rs.standard_normal(size=(J, K)) for a in A]
    lap = 2 * np.eye(J) - np.eye(J, k=1) - np.eye(J, k=-1)
    lap[0, 0] = lap[-1, -1] = 1
    for kw in (dict(non_negative={0: True, 2: True}, generalized_l2_penalty={1: lap}, parafac2=True),
               dict(non_negative={0: True, 1: True}, tv_penalty={2: 0.05}, l1_penalty={2: 0.01}),
               dict(non_negative={0: True, 2: True}, regs=[[], [P.UnitSimplex()], []]),
               dict(non_negative=True, regs=[[], [P.Parafac2(n_iter=2)], []])):
        common = dict(random_state=0, n_iter_max=15, tol=None, absolute_tol=None)
        ref = cmf_aoadmm(X, R, **common, **kw)
        got = cmf_aoadmm([x.astype(np.float32) for x in X], R, **common, **kw)
        for a, b in zip((ref[1][0], np.concatenate(ref[1][1]), ref[1][2]),
                        (got[1][0], np.concatenate(got[1][1]), got[1][2])):
            assert rel(b, a) < 2e-3, (sorted(kw), rel(b, a))


@pytest.mark.parametrize("name", ["init_svd_unconstrained", "init_threshold_svd_nn"])
def test_svd_inits_match_live_oracle(name):
    """init="svd" / "threshold_svd" (decomposition.py:42-53): the start comes from the same host LAPACK call as the
    reference's, the iterations run on the device; trajectory against the oracle run live on this box (the oracle
    itself is pinned against the reference's trajectory for these cases by tests/test_oracle.py), for list input and
    for a device-resident PackedMatrices."""
    from matcouply_b200 import PackedMatrices, cmf_aoadmm
    from oracle import aoadmm_oracle as O

    g, X, rank, kw = load_case(name)
    kw = dict(kw)
    kw.pop("n_iter_max", None)
    for k in (1, 2, 10):
        o = O.ao_admm(X, rank, n_iter_max=k, tol=None, absolute_tol=None, **kw)
        cmf = cmf_aoadmm(X, rank, n_iter_max=k, tol=None, absolute_tol=None, **kw)
        _, (A, B_is, C) = cmf
        assert rel(A, o["A"]) < 1e-8 and rel(C, o["C"]) < 1e-8
        assert rel(np.concatenate(B_is, 0), np.concatenate(o["B_is"], 0)) < 1e-8
    packed = PackedMatrices.from_list(X, torch.float64, torch.device("cuda", 0))
    cmf2 = cmf_aoadmm(packed, rank, n_iter_max=10, tol=None, absolute_tol=None, **kw)
    assert rel(cmf2[1][2], o["C"]) < 1e-8
    with pytest.raises(NotImplementedError):
        cmf_aoadmm(X, rank, init="parafac2_als", n_iter_max=1)


from oracle.random_configs import random_config as _random_config  # noqa: E402  (shared with the CPU pin)


@pytest.mark.parametrize("seed", range(24))
def test_random_penalty_combinations_match_oracle(seed):
    """Differential test over random combinations of ALL penalty kinds, Parafac2 options, feasibility-penalty modes,
    l2 penalties, inner iteration counts and aux/dual initialisations: 10 outer iterations against the oracle."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    X, R, kw = _random_config(seed)
    pf2_zero = kw["aux_init"] == "zeros" and any(p[0] == "Parafac2" for p in kw["regs_spec"][1])
    if pf2_zero:
        kw["aux_init"] = "random_uniform"  # Delta = 0 makes the reference's Procrustes step undefined
    o = O.ao_admm(X, R, **kw)
    cmf, diag = cmf_aoadmm(X, R, return_errors=True, **product_kwargs(kw))
    _, (A, B_is, C) = cmf
    errs = (rel(A, o["A"]), rel(np.concatenate(B_is, 0), np.concatenate(o["B_is"], 0)), rel(C, o["C"]))
    assert max(errs) < 1e-7, (seed, kw["regs_spec"], kw["constant_feasibility_penalty"], errs)
    np.testing.assert_allclose(diag.regularized_loss, o["regularized_loss"], rtol=1e-7)


def test_indefinite_normal_matrix_follows_the_reference_svd_route():
    """No penalty on B and a NEGATIVE l2_penalty there make the B-mode normal matrix `C^T C o a a^T + l2 I`
    indefinite (the smallest eigenvalue of the Gram is below |l2|): Cholesky does not exist, the reference's SVD
    solve (decomposition.py:252-256) still applies the inverse.  Same iterates as the oracle, no NaN."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(8)
    I, K, R = 6, 3, 4  # K < R: C^T C has rank <= 3, so lhs_B = G o aa^T - 0.05 I has a negative eigenvalue
    X = [rs.uniform(size=(J, K)) for J in (7, 9, 8, 12, 6, 10)]
    kw = dict(l2_penalty=[0.0, -0.05, 0.1], non_negative={0: True}, random_state=4, n_iter_max=3, tol=None,
              absolute_tol=None)
    o = O.ao_admm(X, R, **kw)
    cmf = cmf_aoadmm(X, R, **kw)
    _, (A, B_is, C) = cmf
    assert np.all(np.isfinite(C)) and all(np.all(np.isfinite(b)) for b in B_is)
    errs = (rel(A, o["A"]), rel(np.concatenate(B_is, 0), np.concatenate(o["B_is"], 0)), rel(C, o["C"]))
    assert max(errs) < 1e-7, errs


@pytest.mark.parametrize("kw", [dict(parafac2=True), dict(regs_spec=[[], [["Parafac2", {"n_iter": 2}]], []]),
                                dict(regs_spec=[[], [["Parafac2", {"update_coordinate_matrix": False}]], []])])
def test_parafac2_from_a_zero_coordinate_matrix(kw):
    """aux_init="zeros" gives PARAFAC2 the coordinate matrix Delta = 0: the reference's first Procrustes step is then
    the SVD of a zero matrix, for which LAPACK returns identity factors, i.e. P_i = eye(J_i, R) (penalties.py:1233-1235;
    pinned against the live reference by tests/test_oracle.py::test_oracle_vs_live_reference).  The engine takes the
    same step (fixed identity basis for that one update) instead of dropping every direction of the zero Gram."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(21)
    X = [rs.uniform(size=(J, 11)) for J in (8, 13, 9, 10, 21, 7)]
    common = dict(non_negative={0: True, 2: True}, aux_init="zeros", random_state=2, n_iter_max=15, tol=None,
                  absolute_tol=None)
    o = O.ao_admm(X, 3, **common, **kw)
    cmf, admm, diag = cmf_aoadmm(X, 3, return_errors=True, return_admm_vars=True, **common, **product_kwargs(kw))
    _, (A, B_is, C) = cmf
    errs = (rel(A, o["A"]), rel(np.concatenate(B_is, 0), np.concatenate(o["B_is"], 0)), rel(C, o["C"]))
    assert max(errs) < 1e-8, errs
    np.testing.assert_allclose(diag.regularized_loss, o["regularized_loss"], rtol=1e-8)
    bases, delta = admm.auxes[1][0]
    assert rel(delta, o["aux"][1][0][1]) < 1e-7
    assert rel(np.concatenate(bases, 0), np.concatenate(o["aux"][1][0][0], 0)) < 1e-6


@pytest.mark.parametrize("R", [4, 8, 20])
def test_one_array_companion_and_fused_gap_terms_match_the_oracle(R):
    """Non-negativity next to PARAFAC2 at a rank the steady-state row pass serves (fp64, R % 4 == 0): between outer
    iterations the companion lives as ONE array T = x + dual and its feasibility-gap terms come out of the last row
    pass (csrc/pf2_rowpass_v2.cuh).  The reported gaps, the losses and the materialised ADMM variables (aux = prox(T),
    dual = T - aux on demand) must equal the reference algorithm's, also when the stopping rule reads them every
    iteration, and a run that materialises in between (return_admm_vars) must not disturb the iterates."""
    from matcouply_b200 import cmf_aoadmm
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(60 + R)
    I, K = 9, 3 * R + 5
    Js = list(rs.randint(R, 200, size=I - 1)) + [64]
    A, C = rs.uniform(0.2, 1.2, size=(I, R)), rs.uniform(size=(K, R))
    X = [(rs.uniform(size=(J, R)) * a) @ C.T + 0.1 * rs.standard_normal(size=(J, K)) for J, a in zip(Js, A)]
    kw = dict(non_negative=True, parafac2=True, random_state=1, n_iter_max=9, tol=None, absolute_tol=None)
    o = O.ao_admm(X, R, **kw)
    cmf, admm, diag = cmf_aoadmm(X, R, return_errors=True, return_admm_vars=True, **kw)
    _, (Ag, Bg, Cg) = cmf
    assert rel(Ag, o["A"]) < 1e-8 and rel(Cg, o["C"]) < 1e-8
    assert rel(np.concatenate(Bg, 0), np.concatenate(o["B_is"], 0)) < 1e-8
    np.testing.assert_allclose(diag.regularized_loss, o["regularized_loss"], rtol=1e-8)
    gaps = np.array([[v for mode in it for v in mode] for it in diag.feasibility_gaps])
    ogaps = np.array([[v for mode in it for v in mode] for it in o["feasibility_gaps"]])
    np.testing.assert_allclose(gaps, ogaps, rtol=1e-6, atol=1e-13)
    nn_aux, nn_dual = admm.auxes[1][1], admm.duals[1][1]
    assert rel(np.concatenate(nn_aux, 0), np.concatenate(o["aux"][1][1], 0)) < 1e-7
    assert rel(np.concatenate(nn_dual, 0), np.concatenate(o["dual"][1][1], 0)) < 1e-6
    assert min(a.min() for a in nn_aux) >= 0
    # one- and two-iteration runs (first B-update starts from explicit state, the second from T) also agree
    for k in (1, 2):
        ok = O.ao_admm(X, R, **dict(kw, n_iter_max=k))
        ck, ak = cmf_aoadmm(X, R, return_admm_vars=True, **dict(kw, n_iter_max=k))
        assert rel(np.concatenate(ck[1][1], 0), np.concatenate(ok["B_is"], 0)) < 1e-9
        assert rel(np.concatenate(ak.duals[1][1], 0), np.concatenate(ok["dual"][1][1], 0)) < 1e-8


@pytest.mark.parametrize("kw", [
    dict(non_negative=True),
    dict(non_negative=True, l1_penalty={1: 0.05, 2: 0.02}),
    dict(non_negative={0: True, 2: True}),                       # no penalty on the B-mode
    dict(lower_bound={1: -0.2}, upper_bound={1: 0.7}, l2_penalty=0.01),
    dict(non_negative=True, l2_norm_bound={0: 1.0}, constant_feasibility_penalty=True),
])
def test_single_read_fused_pass_matches_two_pass_schedule(kw, monkeypatch):
    """with row-local B-mode penalties the engine runs ONE pass over X per outer iteration
    (csrc/xfused.cu).  Same iterates, ADMM variables and diagnostics as the two-pass schedule (round-off apart: the
    products are summed in another order)."""
    from matcouply_b200 import _engine, _ops, cmf_aoadmm

    rs = np.random.RandomState(5)
    I, K, R = 23, 70, 5
    Js = list(rs.randint(4 * R, 160, size=I - 1)) + [64]
    X = [rs.uniform(size=(J, K)) for J in Js]
    calls = {"fused": 0, "y": 0}
    real_fused, real_y = _ops.xstream_fused_local, _ops.xstream_y
    monkeypatch.setattr(_ops, "xstream_fused_local", lambda *a, **k: (calls.__setitem__("fused", calls["fused"] + 1),
                                                                       real_fused(*a, **k))[1])
    monkeypatch.setattr(_ops, "xstream_y", lambda *a, **k: (calls.__setitem__("y", calls["y"] + 1), real_y(*a, **k))[1])
    out = []
    default = _engine.FUSION_DEFAULTS["x1"]
    for x1 in (True, False):
        _engine.FUSION_DEFAULTS.update(x1=x1)
        calls.update(fused=0, y=0)
        try:
            cmf, admm, diag = cmf_aoadmm(X, R, n_iter_max=8, tol=None, absolute_tol=None, random_state=3,
                                         return_admm_vars=True, return_errors=True, **kw)
        finally:
            _engine.FUSION_DEFAULTS.update(x1=default)
        # fused: one pass per iteration (+ the Y pass of the initial fit); two-pass: no fused launch at all
        assert (calls["fused"], calls["y"]) == ((8, 1) if x1 else (0, 9))
        out.append((cmf, admm, diag))
    (c1, a1, d1), (c2, a2, d2) = out
    tol = 1e-9
    assert rel(c1[1][0], c2[1][0]) < tol and rel(c1[1][2], c2[1][2]) < tol
    assert rel(np.concatenate(c1[1][1], 0), np.concatenate(c2[1][1], 0)) < tol
    np.testing.assert_allclose(d1.regularized_loss, d2.regularized_loss, rtol=1e-9)
    np.testing.assert_allclose(d1.rec_errors, d2.rec_errors, rtol=1e-9)
    for m in range(3):
        for x1_, x2_ in zip(a1.duals[m] + a1.auxes[m], a2.duals[m] + a2.auxes[m]):
            assert rel(np.concatenate(x1_, 0) if m == 1 else x1_, np.concatenate(x2_, 0) if m == 1 else x2_) < 10 * tol
