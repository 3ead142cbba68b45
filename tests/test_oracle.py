"""CPU tests that PIN the oracle: against the golden vectors generated from the unmodified reference
(oracle/gen_golden.py), against scikit-learn for the isotonic building block, and — when /root/reference is mounted
(build container only) — against the reference run live."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import aoadmm_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(HERE, "golden", "traj_*.npz")))


@pytest.fixture(scope="session", autouse=True)
def build_oracle_c():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def load_case(name):
    g = np.load(os.path.join(HERE, "golden", f"traj_{name}.npz"))
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


@pytest.mark.parametrize("name", [c for c in CASES if c != "c0_readme"])
def test_oracle_reproduces_reference_goldens(name):
    g, X, rank, kw = load_case(name)
    traj = []
    o = O.ao_admm(X, rank, trajectory=traj, **kw)
    assert o["n_iter"] == int(g["n_iter"]) and o["message"] == str(g["message"])
    np.testing.assert_allclose(o["regularized_loss"], g["regularized_loss"], rtol=1e-11)
    np.testing.assert_allclose(o["rec_errors"], g["rec_errors"], rtol=1e-10)
    for k in range(g["A_traj"].shape[0] if g["A_traj"].ndim == 3 else 0):
        np.testing.assert_allclose(traj[k]["A"], g["A_traj"][k], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(np.concatenate(traj[k]["B_is"], 0), g["B_traj"][k], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(traj[k]["C"], g["C_traj"][k], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(o["A"], g["A"], rtol=1e-8, atol=1e-12)


def test_oracle_readme_first_iterations():
    g, X, rank, kw = load_case("c0_readme")
    traj = []
    o = O.ao_admm(X, rank, trajectory=traj, n_iter_max=50, **kw)
    np.testing.assert_allclose(o["regularized_loss"], g["regularized_loss"][:51], rtol=1e-11)
    for k in range(50):
        np.testing.assert_allclose(np.concatenate(traj[k]["B_is"], 0), g["B_traj"][k], rtol=1e-9, atol=1e-13)


def test_unimodal_oracle_c_and_python_twin_bit_exact():
    g = np.load(os.path.join(HERE, "golden", "operators.npz"))
    assert O._load_c(), "oracle/_build/liboracle.so missing"
    for k in range(int(g["uni_count"])):
        y, fit, peaks, nn = g[f"uni{k}_y"], g[f"uni{k}_fit"], g[f"uni{k}_peaks"], bool(g[f"uni{k}_nn"])
        f1, p1 = O.unimodal_regression(y, nn, return_peaks=True)
        assert np.array_equal(f1, fit) and np.array_equal(p1, peaks)
        if y.shape[0] <= 60:
            f2, p2 = O.unimodal_regression(y, nn, return_peaks=True, force_python=True)
            assert np.array_equal(f2, fit) and np.array_equal(p2, peaks)


def test_prefix_isotonic_vs_sklearn():
    """Every prefix of the PAVA fit equals scikit-learn's isotonic regression of that prefix
    (same check as the reference's tests/test_unimodal_regression.py:21-41)."""
    from sklearn.isotonic import IsotonicRegression

    rs = np.random.RandomState(0)
    for n in (5, 31, 120):
        for nn in (False, True):
            y = rs.standard_normal(n)
            level, start, err = O._prefix_isotonic_py(y, nn)
            for end in range(1, n + 1):
                fit = O._expand_blocks(end, level, start)
                sk = IsotonicRegression(y_min=0 if nn else None).fit_transform(np.arange(end), y[:end])
                np.testing.assert_allclose(fit, sk, atol=1e-12)
                assert abs(err[end] - np.sum((fit - y[:end]) ** 2)) < 1e-9


def test_tv_oracle_kkt():
    """The TV prox is the unique minimiser of 0.5||x - y||^2 + lam TV(x).  KKT certificate: with u = cumsum(y - x),
    u[-1] = 0, |u[k]| <= lam and u[k] = -lam sign(x[k+1] - x[k]) wherever x jumps.  This pins the oracle's restatement
    of Condat's algorithm without the (un-installable) condat_tv package."""
    rs = np.random.RandomState(0)
    for n in (1, 2, 3, 5, 10, 50, 200, 1000):
        for lam in (1e-3, 0.1, 1.0, 10.0):
            y = rs.standard_normal(n) * rs.choice([0.1, 1, 5])
            if n > 20:
                y += np.repeat(rs.standard_normal(n // 10 + 1), 10)[:n] * 3
            x = O.tv_denoise_1d(y, lam)
            u = np.cumsum(y - x)
            assert abs(u[-1]) < 1e-9 * max(1.0, np.abs(y).sum())
            assert np.all(np.abs(u[:-1]) <= lam * (1 + 1e-9) + 1e-12)
            d = np.diff(x)
            jump = np.abs(d) > 1e-12
            np.testing.assert_allclose(u[:-1][jump], -lam * np.sign(d[jump]), atol=1e-9)


def test_simplex_oracle_vs_sort_projection():
    """The bisection of penalties.py:941-969 against the exact sort-based projection onto the unit simplex."""
    rs = np.random.RandomState(1)
    for n in (1, 2, 5, 30, 400):
        M = rs.standard_normal((n, 4)) * 2
        out = O.UnitSimplexP().prox(M, 1.0, None)
        for r in range(4):
            y = np.sort(M[:, r])[::-1]
            css = np.cumsum(y) - 1
            k = np.arange(1, n + 1)
            ok = y - css / k > 0
            mu = css[ok][-1] / k[ok][-1]
            assert np.abs(out[:, r] - np.clip(M[:, r] - mu, 0, None)).max() < 1e-11
        np.testing.assert_allclose(out.sum(0), 1.0, atol=1e-10)


def test_next_penalty_operator_goldens():
    """GeneralizedL2 / UnitSimplex (reference classes) and TV (reference class on the condat_tv stand-in)."""
    g = np.load(os.path.join(HERE, "golden", "operators.npz"))
    M = g["prox_in"]
    np.testing.assert_allclose(O.GeneralizedL2P(g["gl2_matrix"]).prox(M, 1.3, None), g["prox_gl2"], rtol=1e-12)
    np.testing.assert_allclose(O.GeneralizedL2P(g["gl2_matrix"]).value(M), float(g["gl2_value"]), rtol=1e-12)
    np.testing.assert_allclose(O.UnitSimplexP().prox(M, 1.3, None), g["prox_simplex"], atol=1e-14)
    np.testing.assert_allclose(O.TotalVariationP(0.3).prox(M, 1.3, None), g["prox_tv"], atol=1e-14)
    np.testing.assert_allclose(O.TotalVariationP(0.3, 0.2).value(M), float(g["tv_value"]), rtol=1e-12)
    np.testing.assert_allclose(O.UnitSimplexP().prox(g["prox_in_long"], 0.7, None), g["prox_simplex_long"], atol=1e-14)


def test_closed_form_l2_update():
    """With l2_penalty=1 and no constraint one C-update equals solve(Gram + I, rhs)
    (reference tests/test_decomposition.py:1423-1443)."""
    rs = np.random.RandomState(2)
    A, C = rs.uniform(size=(5, 3)), rs.uniform(size=(8, 3))
    Bs = [rs.uniform(size=(J, 3)) for J in (6, 7, 9, 4, 5)]
    X = [(B * a) @ C.T for B, a in zip(Bs, A)]
    Cn, _, _ = O.solve_mode_C(X, [], A, Bs, C, [], [], 1.0, 5, 1)
    lhs = sum((B * a).T @ (B * a) for B, a in zip(Bs, A)) + np.eye(3)
    rhs = sum(x.T @ (B * a) for x, B, a in zip(X, Bs, A))
    np.testing.assert_allclose(Cn, np.linalg.solve(lhs, rhs.T).T, rtol=1e-10)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/matcouply"), reason="reference only in the build container")
def test_oracle_vs_live_reference():
    code = r"""
import sys, os
sys.dont_write_bytecode = True
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache_test")
sys.path[:0] = [%r, "/root/reference/src", %r]
import numpy as np
from matcouply.decomposition import cmf_aoadmm
from oracle.aoadmm_oracle import ao_admm
rs = np.random.RandomState(9)
X = [rs.uniform(size=(J, 11)) for J in (8, 13, 9, 10, 21, 7)]
kw = dict(non_negative=True, parafac2=True, unimodal={1: True}, l1_penalty={2: 0.05}, l2_norm_bound=[0.0, 2.0, 0],
          random_state=4, n_iter_max=40)
cmf, d = cmf_aoadmm(X, 3, return_errors=True, **kw)
o = ao_admm(X, 3, **kw)
assert d.n_iter == o["n_iter"]
np.testing.assert_allclose(o["regularized_loss"], d.regularized_loss, rtol=1e-11)
np.testing.assert_allclose(o["A"], cmf[1][0], rtol=1e-9, atol=1e-13)
# PARAFAC2 from a ZERO coordinate matrix (aux_init="zeros"): the reference's first Procrustes step is the SVD of a zero
# matrix, whose LAPACK factors are identities (P_i = eye(J_i, R)); the oracle must take the same path
kw = dict(parafac2=True, non_negative={0: True, 2: True}, aux_init="zeros", random_state=2, n_iter_max=15, tol=None,
          absolute_tol=None)
cmf, d = cmf_aoadmm(X, 3, return_errors=True, **kw)
o = ao_admm(X, 3, **kw)
np.testing.assert_allclose(o["regularized_loss"], d.regularized_loss, rtol=1e-11)
np.testing.assert_allclose(o["C"], cmf[1][2], rtol=1e-9, atol=1e-13)
print("OK")
""" % (os.path.join(ROOT, "oracle", "_tl_standin"), ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("kw", [dict(non_negative=True),
                                dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}),
                                dict(parafac2=True, l1_penalty={2: 0.05})])
def test_torch_cpu_oracle_matches_numpy_oracle(kw):
    """The torch-CPU column of bench.py (oracle/aoadmm_torch_cpu.py: the reference under TensorLy's PyTorch backend,
    SURVEY.md §8c) follows the pinned NumPy oracle: to round-off in float64, at single precision in float32 (the
    backend's default dtype; the reference's own torch tests scale their tolerances by 500, tests/utils.py:6-9)."""
    import torch

    from oracle.aoadmm_torch_cpu import ao_admm_torch_cpu

    rs = np.random.RandomState(3)
    R = 3
    X = [rs.uniform(size=(J, R)) @ rs.uniform(size=(R, 14)) + 0.1 * rs.standard_normal((J, 14))
         for J in rs.randint(6, 20, size=9)]
    o = O.ao_admm(X, R, n_iter_max=6, random_state=0, tol=None, absolute_tol=None, **kw)
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 2e-4)):
        t = ao_admm_torch_cpu(X, R, n_iter_max=6, random_state=0, dtype=dtype, **kw)
        np.testing.assert_allclose(t["A"], o["A"], rtol=tol, atol=tol)
        np.testing.assert_allclose(t["C"], o["C"], rtol=tol, atol=tol)
        for b, ob in zip(t["B_is"], o["B_is"]):
            np.testing.assert_allclose(b, ob, rtol=tol, atol=tol)
        np.testing.assert_allclose(t["regularized_loss"], o["regularized_loss"][1:], rtol=tol)
        np.testing.assert_allclose(t["rec_errors"], o["rec_errors"][1:], rtol=tol)


def test_torch_cpu_oracle_refuses_numpy_only_penalties():
    from oracle.aoadmm_torch_cpu import ao_admm_torch_cpu

    with pytest.raises(TypeError):  # unimodal is not even a keyword of the torch column (penalties.py:1008-1009)
        ao_admm_torch_cpu([np.ones((4, 3))], 2, unimodal={1: True})


@pytest.mark.parametrize("seed", range(24))
def test_oracle_matches_reference_on_random_configs(seed):
    """The 24 seeded random combinations of all penalty kinds / Parafac2 options / feasibility-penalty modes / inits of
    the GPU differential test (tests/test_gpu_aoadmm.py::test_random_penalty_combinations_match_oracle), run through
    the UNMODIFIED reference by oracle/gen_golden_random.py: the oracle reproduces the reference's final factors and
    its whole loss / error history, which closes the chain CUDA path == oracle == reference on these problems."""
    from oracle.random_configs import random_config, reference_safe

    g = np.load(os.path.join(HERE, "golden", "random_configs.npz"))
    X, R, kw = random_config(seed)
    o = O.ao_admm(X, R, **reference_safe(kw))
    np.testing.assert_allclose(o["regularized_loss"], g[f"s{seed}_loss"], rtol=1e-9)
    np.testing.assert_allclose(o["rec_errors"], g[f"s{seed}_rec"], rtol=1e-9)
    for got, key in ((o["A"], "A"), (np.concatenate(o["B_is"], 0), "B"), (o["C"], "C")):
        ref = g[f"s{seed}_{key}"]
        assert np.linalg.norm(got - ref) <= 1e-9 * np.linalg.norm(ref), (seed, key)
