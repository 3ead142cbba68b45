"""The reference's own known-answer tests for the hot path (reference tests/test_decomposition.py; SURVEY.md §4 "what pins
the hot path"), written against ``matcouply_b200`` — no oracle involved, only closed forms and exact data:

* fit term equals the injected noise (:122-133) and the naive sum of squared residuals along a run (:1961-2010),
* first loss value (:1844-1874), l2 penalty included (:1707-1755),
* closed-form regularised least-squares updates of A / B_i / C with ``l2_penalty=1`` (:1080-1100, 1339-1361, 1423-1443),
* exact-data recovery (:1446-1565), PARAFAC2 makes the non-negative CMF unique (:1597-1625),
* stopping information (:1758-1841), zero iterations (:1877-1889), feasibility info without tolerances (:1892-1899),
* frozen modes (:1628-1693), ``regs`` not mutated (:1902-1921), ``constant_feasibility_penalty`` validation (:1924-1958).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ragged_cmf(seed, rank=3, I=5, K=6, Jlo=5, Jhi=11):
    rs = np.random.RandomState(seed)
    Js = [int(j) for j in rs.randint(Jlo, Jhi, size=I)]
    A = rs.uniform(0.1, 1.1, size=(I, rank))
    B_is = [rs.uniform(size=(J, rank)) for J in Js]
    C = rs.uniform(size=(K, rank))
    return rs, A, B_is, C, [(B * a) @ C.T for B, a in zip(B_is, A)]


def _congruence(X, Y):
    """Best-permutation mean absolute column congruence (tensorly.metrics.congruence_coefficient, absolute_value=True)."""
    from itertools import permutations

    Xn, Yn = X / np.linalg.norm(X, axis=0), Y / np.linalg.norm(Y, axis=0)
    M = np.abs(Xn.T @ Yn)
    return max(np.mean([M[i, p[i]] for i in range(M.shape[0])]) for p in permutations(range(M.shape[1])))


def test_reconstruction_error_equals_noise_and_naive_sse():
    from matcouply_b200 import cmf_aoadmm

    rs, A, B_is, C, mats = _ragged_cmf(0)
    noise = [0.1 * rs.standard_normal(size=m.shape) for m in mats]
    noisy = [m + n for m, n in zip(mats, noise)]
    normX = np.sqrt(sum(np.sum(m ** 2) for m in noisy))
    _, diag = cmf_aoadmm(noisy, 3, init=(None, (A, B_is, C)), n_iter_max=0, return_errors=True)
    np.testing.assert_allclose(diag.rec_errors[0] * normX, np.sqrt(sum(np.sum(n ** 2) for n in noise)), rtol=1e-10)
    # along a run the expanded fit formula (no X read) coincides with the naive residual of the returned factors
    for k in (1, 7, 30):
        cmf, diag = cmf_aoadmm(noisy, 3, n_iter_max=k, non_negative=True, random_state=1, tol=None, absolute_tol=None,
                               return_errors=True)
        naive = np.sqrt(sum(np.sum((x - xh) ** 2) for x, xh in zip(noisy, cmf.to_matrices()))) / normX
        np.testing.assert_allclose(diag.rec_errors[-1], naive, rtol=1e-5, atol=1e-7)


def test_first_loss_value_and_l2_penalty_included():
    from matcouply_b200 import cmf_aoadmm

    rs, A, B_is, C, mats = _ragged_cmf(1)
    A2, C2 = A + 0.1, C + 0.2  # start away from the truth
    B2 = [b + 0.05 for b in B_is]
    normX2 = sum(np.sum(m ** 2) for m in mats)
    sse = sum(np.sum((m - (b * a) @ C2.T) ** 2) for m, b, a in zip(mats, B2, A2))
    l2 = [0.3, 0.2, 0.1]
    gamma = 0.7
    _, diag = cmf_aoadmm(mats, 3, init=(None, (A2, B2, C2)), n_iter_max=0, return_errors=True, l2_penalty=l2,
                         l1_penalty={0: gamma})
    want = 0.5 * sse / normX2
    want += 0.5 * (l2[0] * np.sum(A2 ** 2) + l2[1] * sum(np.sum(b ** 2) for b in B2) + l2[2] * np.sum(C2 ** 2))
    want += gamma * np.sum(np.abs(A2))
    np.testing.assert_allclose(diag.regularized_loss[0], want, rtol=1e-12)
    _, d0 = cmf_aoadmm(mats, 3, init=(None, (A2, B2, C2)), n_iter_max=0, return_errors=True)
    np.testing.assert_allclose(d0.regularized_loss[0], 0.5 * sse / normX2, rtol=1e-12)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_closed_form_update_with_l2_penalty(mode):
    """With only mode `mode` updated, no constraint and l2_penalty=1 on it, one outer iteration is the regularised
    least-squares solution solve(Gram + I, rhs) (the l2 branch of test_admm_update_A/B/C)."""
    from matcouply_b200 import cmf_aoadmm

    rs, A, B_is, C, mats = _ragged_cmf(2 + mode)
    A0, C0 = rs.uniform(size=A.shape), rs.uniform(size=C.shape)
    B0 = [rs.uniform(size=b.shape) for b in B_is]
    l2 = [0, 0, 0]
    l2[mode] = 1
    upd = dict(update_A=mode == 0, update_B_is=mode == 1, update_C=mode == 2)
    cmf = cmf_aoadmm(mats, 3, init=(None, (A0, B0, C0)), n_iter_max=1, l2_penalty=l2, tol=None, absolute_tol=None, **upd)
    _, (A1, B1, C1) = cmf
    if mode == 2:
        lhs = sum((b * a).T @ (b * a) for b, a in zip(B0, A0)) + np.eye(3)
        rhs = sum(x.T @ (b * a) for x, b, a in zip(mats, B0, A0))
        np.testing.assert_allclose(C1, np.linalg.solve(lhs, rhs.T).T, rtol=1e-9)
        np.testing.assert_array_equal(A1, A0)
    elif mode == 1:
        for x, a, b1 in zip(mats, A0, B1):
            lhs = (C0 * a).T @ (C0 * a) + np.eye(3)
            np.testing.assert_allclose(b1, np.linalg.solve(lhs, (x @ (C0 * a)).T).T, rtol=1e-9)
        np.testing.assert_array_equal(C1, C0)
    else:
        for x, b, a1 in zip(mats, B0, A1):
            lhs = (b.T @ b) * (C0.T @ C0) + np.eye(3)
            np.testing.assert_allclose(a1, np.linalg.solve(lhs, np.diag(b.T @ x @ C0)), rtol=1e-9)
        np.testing.assert_array_equal(C1, C0)


@pytest.mark.parametrize("nn", [False, True])
def test_exact_data_is_fitted(nn):
    from matcouply_b200 import cmf_aoadmm

    _, A, B_is, C, mats = _ragged_cmf(7, rank=2, I=6, K=7)
    cmf, diag = cmf_aoadmm(mats, 2, n_iter_max=5000, non_negative=nn, return_errors=True, random_state=0, tol=1e-12,
                           absolute_tol=1e-12)
    assert diag.rec_errors[-1] < 1e-2
    weights, (Ah, Bh, Ch) = cmf
    assert weights is None and Ah.shape == A.shape and Ch.shape == C.shape
    assert [b.shape for b in Bh] == [b.shape for b in B_is]
    if nn:
        assert Ah.min() >= -1e-4 and Ch.min() >= -1e-4  # primal factors are feasible up to the feasibility gap


def test_parafac2_makes_nn_cmf_unique():
    from matcouply_b200 import cmf_aoadmm

    rng = np.random.RandomState(1)
    rank = 2
    A = rng.uniform(0.1, 1.1, size=(10, rank))
    B_0 = rng.uniform(0, 1, size=(7, rank))
    B_is = [np.roll(B_0, i, axis=1) for i in range(10)]
    C = rng.uniform(0, 1, size=(10, rank))
    mats = [(b * a) @ C.T for b, a in zip(B_is, A)]
    best, best_cmf = [float("inf")], None
    for init in range(5):
        out, diag = cmf_aoadmm(mats, rank, n_iter_max=1000, return_errors=True, non_negative=[True, True, True],
                               parafac2=True, random_state=init)
        if diag.regularized_loss[-1] < best[-1] and diag.satisfied_feasibility_condition:
            best, best_cmf = diag.regularized_loss, out
    assert best[-1] < 1e-5
    assert _congruence(A, best_cmf[1][0]) > 0.95
    assert _congruence(C, best_cmf[1][2]) > 0.95
    for b, bh in zip(B_is, best_cmf[1][1]):
        assert _congruence(b, bh) > 0.95


def test_stopping_information():
    from matcouply_b200 import cmf_aoadmm

    _, _, _, _, mats = _ragged_cmf(9)
    inf = float("inf")
    n = 10
    expect = [
        ((-inf, -inf, -inf), False, False, "MAXIMUM NUMBER OF ITERATIONS REACHED", n + 1),
        ((-inf, -inf, inf), False, True, "MAXIMUM NUMBER OF ITERATIONS REACHED", n + 1),
        ((inf, -inf, inf), True, True, "FEASIBILITY GAP CRITERION AND RELATIVE LOSS CRITERION SATISFIED", 2),
        ((-inf, inf, inf), True, True, "FEASIBILITY GAP CRITERION AND ABSOLUTE LOSS CRITERION SATISFIED", 2),
    ]
    for (tol, atol, ftol), stop, feas, msg, length in expect:
        _, d = cmf_aoadmm(mats, 3, n_iter_max=n, return_errors=True, tol=tol, absolute_tol=atol, feasibility_tol=ftol,
                          non_negative=True, random_state=0)
        assert bool(d.satisfied_stopping_condition) == stop and bool(d.satisfied_feasibility_condition) == feas
        assert d.message == msg
        assert len(d.regularized_loss) == len(d.rec_errors) == len(d.feasibility_gaps) == length
    # zero iterations; feasibility information without any tolerance
    out, d = cmf_aoadmm(mats, 3, n_iter_max=0, return_errors=True, random_state=0)
    assert d.n_iter == 0 and len(d.rec_errors) == 1
    _, d = cmf_aoadmm(mats, 3, n_iter_max=5, return_errors=True, tol=None, absolute_tol=None, non_negative=True,
                      random_state=0)
    assert d.satisfied_stopping_condition is None and len(d.feasibility_gaps) == 6
    _, d = cmf_aoadmm(mats, 3, n_iter_max=5, return_errors=True, feasibility_tol=None, random_state=0)
    assert d.satisfied_feasibility_condition is None


def test_frozen_modes_regs_untouched_and_validation():
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.penalties import NonNegativity

    rs, A, B_is, C, mats = _ragged_cmf(11)
    init = (None, (A + 0.3, [b + 0.1 for b in B_is], C + 0.2))
    for frozen in ("update_A", "update_B_is", "update_C"):
        cmf = cmf_aoadmm(mats, 3, init=init, n_iter_max=5, non_negative=True, **{frozen: False})
        _, (A1, B1, C1) = cmf
        if frozen == "update_A":
            np.testing.assert_array_equal(A1, init[1][0])
            assert not np.allclose(C1, init[1][2])
        elif frozen == "update_C":
            np.testing.assert_array_equal(C1, init[1][2])
            assert not np.allclose(A1, init[1][0])
        else:
            for b, b0 in zip(B1, init[1][1]):
                np.testing.assert_array_equal(b, b0)
    regs = [[NonNegativity()], [NonNegativity(), NonNegativity()], []]
    lens = [len(r) for r in regs]
    cmf_aoadmm(mats, 3, n_iter_max=2, regs=regs, non_negative=True, l1_penalty={2: 0.1})
    assert [len(r) for r in regs] == lens
    for bad in ("C", "both", ""):
        if bad == "":
            continue
        with pytest.raises(ValueError):
            cmf_aoadmm(mats, 3, n_iter_max=1, constant_feasibility_penalty=bad)
    for ok in (True, False, "A", "B"):
        cmf_aoadmm(mats, 3, n_iter_max=1, constant_feasibility_penalty=ok, non_negative=True)
    # matrix-wise penalties on mode 0 need a constant feasibility penalty, as in the reference (README.rst:88)
    with pytest.raises(AttributeError):
        cmf_aoadmm(mats, 3, n_iter_max=1, l2_norm_bound={0: 1.0})
    cmf_aoadmm(mats, 3, n_iter_max=1, l2_norm_bound={0: 1.0}, constant_feasibility_penalty="A")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_admm_update_functions_like_the_reference_tests(mode):
    """admm_update_A / _B / _C called directly with the reference's signature (tests/test_decomposition.py:1018-1100,
    1263-1361, 1364-1443): (i) exact data + non-negativity + many inner iterations recover the true factor,
    (ii) one call with random ADMM state equals the oracle's sub-solver, (iii) feasibility gaps of the returned aux."""
    from matcouply_b200 import decomposition as D
    from matcouply_b200.penalties import L1Penalty, NonNegativity
    from oracle import aoadmm_oracle as O

    rs, A, B_is, C, mats = _ragged_cmf(20 + mode, rank=3, I=6, K=7)
    fun = (D.admm_update_A, D.admm_update_B, D.admm_update_C)[mode]
    shapes_aux = (A.shape, None, C.shape)

    def rnd_state():
        if mode == 1:
            return [rs.uniform(size=b.shape) for b in B_is], [rs.uniform(size=b.shape) for b in B_is]
        return rs.uniform(size=shapes_aux[mode]), rs.uniform(size=shapes_aux[mode])

    # (i) recovery from a perturbed start, everything else at the truth
    start = [A.copy(), [b.copy() for b in B_is], C.copy()]
    if mode == 1:
        start[1] = [rs.uniform(size=b.shape) for b in B_is]
    else:
        start[mode] = rs.uniform(size=start[mode].shape)
    aux, dual = rnd_state()
    args = (mats, [NonNegativity()], (None, start), [aux], [dual], 0, 2000, None, 1)
    out = fun(*args, False, None) if mode != 2 else fun(*args, None)
    got = out[0][1][mode]
    if mode == 1:
        for b, bt in zip(got, B_is):
            np.testing.assert_allclose(b, bt, rtol=1e-5, atol=1e-7)
    else:
        np.testing.assert_allclose(got, (A, None, C)[mode], rtol=1e-5, atol=1e-7)
    # (ii) five inner iterations from a random state against the oracle's sub-solver, two penalties
    start = [rs.uniform(size=A.shape), [rs.uniform(size=b.shape) for b in B_is], rs.uniform(size=C.shape)]
    a1, d1 = rnd_state()
    a2, d2 = rnd_state()
    cp = lambda v: [x.copy() for x in v] if isinstance(v, list) else v.copy()  # noqa: E731
    regs_p = [NonNegativity(), L1Penalty(0.05)]
    regs_o = [O.NonNeg(), O.L1P(0.05)]
    args = (mats, regs_p, (None, [start[0].copy(), cp(start[1]), start[2].copy()]), [cp(a1), cp(a2)], [cp(d1), cp(d2)],
            0.1, 5, None, 1.5)
    out = fun(*args, False, None) if mode != 2 else fun(*args, None)
    if mode == 1:
        ref, raux, rdual = O.solve_mode_B(mats, regs_o, start[0], cp(start[1]), start[2], [cp(a1), cp(a2)],
                                          [cp(d1), cp(d2)], 0.1, 5, 1.5, False)
        np.testing.assert_allclose(np.concatenate(out[0][1][1]), np.concatenate(ref), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.concatenate(out[1][1]), np.concatenate(raux[1]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.concatenate(out[2][0]), np.concatenate(rdual[0]), rtol=1e-9, atol=1e-12)
    elif mode == 2:
        ref, raux, rdual = O.solve_mode_C(mats, regs_o, start[0], start[1], start[2].copy(), [a1.copy(), a2.copy()],
                                          [d1.copy(), d2.copy()], 0.1, 5, 1.5)
        np.testing.assert_allclose(out[0][1][2], ref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out[1][1], raux[1], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out[2][0], rdual[0], rtol=1e-9, atol=1e-12)
    else:
        ref, raux, rdual, inter = O.solve_mode_A(mats, regs_o, start[0].copy(), start[1], start[2], [a1.copy(), a2.copy()],
                                                 [d1.copy(), d2.copy()], 0.1, 5, 1.5, False)
        np.testing.assert_allclose(out[0][1][0], ref, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out[1][1], raux[1], rtol=1e-9, atol=1e-12)
        rhses, cross = out[3]
        np.testing.assert_allclose(np.stack(rhses), np.stack(inter[0]), rtol=1e-10)
        np.testing.assert_allclose(np.stack(cross), np.stack(inter[1]), rtol=1e-10)
    # fit term helper (decomposition.py:420-452)
    noisy = [m + 0.05 * rs.standard_normal(size=m.shape) for m in mats]
    naive = np.sqrt(sum(np.sum((x - (b * a) @ C.T) ** 2) for x, b, a in zip(noisy, B_is, A)))
    np.testing.assert_allclose(D._cmf_reconstruction_error(noisy, (None, (A, B_is, C))), naive, rtol=1e-9)


def test_custom_penalty_example_imports_work():
    """The import lines of the reference's custom-penalty example (examples/plot_custom_penalty.py:213-231) work with
    the package name changed: `_doc_utils.copy_ancestor_docstring`, `_unimodal_regression.unimodal_regression`
    (vector, matrix and 3-D input, reference tests/test_unimodal_regression.py:44-115), and a penalty built from them
    runs through `cmf_aoadmm` via `regs` and stays unimodal in all but the last column."""
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200._doc_utils import copy_ancestor_docstring
    from matcouply_b200._unimodal_regression import unimodal_regression
    from matcouply_b200.penalties import HardConstraintMixin, MatrixPenalty
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(0)
    for nn in (False, True):
        Y = rs.standard_normal(size=(40, 5))
        want = O.unimodal_regression(Y, non_negativity=nn)
        np.testing.assert_allclose(unimodal_regression(Y, non_negativity=nn), want, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(unimodal_regression(Y[:, 2], non_negativity=nn), want[:, 2], rtol=1e-12, atol=1e-13)
        T = rs.standard_normal(size=(30, 3, 2))
        got = unimodal_regression(T, non_negativity=nn)
        assert got.shape == T.shape
        np.testing.assert_allclose(got.reshape(30, 6), O.unimodal_regression(T.reshape(30, 6), non_negativity=nn),
                                   rtol=1e-12, atol=1e-13)
    up = np.arange(10.0)
    np.testing.assert_allclose(unimodal_regression(up), up)  # a monotone vector is unimodal already

    class UnimodalAllExceptLast(HardConstraintMixin, MatrixPenalty):
        def __init__(self, non_negativity=False, aux_init="random_uniform", dual_init="random_uniform"):
            super().__init__(aux_init, dual_init)
            self.non_negativity = non_negativity

        @copy_ancestor_docstring
        def factor_matrix_update(self, factor_matrix, feasibility_penalty, aux):
            new = np.copy(factor_matrix)
            new[:, :-1] = unimodal_regression(factor_matrix[:, :-1], non_negativity=self.non_negativity)
            if self.non_negativity:
                new = np.clip(new, 0, None)
            return new

    _, A, B_is, C, mats = _ragged_cmf(3)
    pen = UnimodalAllExceptLast(non_negativity=True)
    cmf, admm = cmf_aoadmm(mats, 3, regs=[[], [pen], []], n_iter_max=5, random_state=0, return_admm_vars=True)
    for aux in admm[0][1][0]:
        body = np.asarray(aux)[:, :-1]
        np.testing.assert_allclose(body, O.unimodal_regression(body, non_negativity=True), rtol=1e-10, atol=1e-12)
