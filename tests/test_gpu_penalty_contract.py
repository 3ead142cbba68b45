"""The reference's penalty contract (src/matcouply/testing/admm_penalty.py:416-551, instantiated per class in
tests/test_penalties.py:66-520), run against the CUDA-backed penalty classes: an invariant point stays, a non-invariant
point moves, the penalty value does not increase, hard constraints report a zero penalty, aux/dual initialisation
shapes and errors (:54-385, condensed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _lap(n):
    L = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    L[0, 0] = L[-1, -1] = 1
    return L


def _cases():
    from matcouply_b200 import penalties as P

    rs = np.random.RandomState(0)
    J, R = 9, 3

    def unimodal(shape):
        t = np.linspace(-1, 1, shape[0])[:, None]
        return np.exp(-(t - rs.uniform(-0.5, 0.5, size=(1, shape[1]))) ** 2 * 8)

    def simplex(shape):
        m = rs.uniform(0.1, 1, size=shape)
        return m / m.sum(0, keepdims=True)

    return [
        # (name, factory, invariant matrix generator, non-invariant generator, rows fixed?)
        ("NonNegativity", lambda: P.NonNegativity(), lambda s: rs.uniform(size=s), lambda s: -rs.uniform(0.1, 1, size=s)),
        ("Box", lambda: P.Box(0.2, 0.7), lambda s: rs.uniform(0.2, 0.7, size=s), lambda s: rs.uniform(1, 2, size=s)),
        ("L1Penalty", lambda: P.L1Penalty(0.1), lambda s: np.zeros(s), lambda s: rs.uniform(1, 2, size=s)),
        ("L1Penalty_nn", lambda: P.L1Penalty(0.1, non_negativity=True), lambda s: np.zeros(s),
         lambda s: rs.uniform(1, 2, size=s)),
        ("L2Ball", lambda: P.L2Ball(1.0), lambda s: rs.uniform(size=s) / (2 * np.sqrt(s[0])),
         lambda s: rs.uniform(2, 3, size=s)),
        ("L2Ball_nn", lambda: P.L2Ball(1.0, non_negativity=True), lambda s: rs.uniform(size=s) / (2 * np.sqrt(s[0])),
         lambda s: -rs.uniform(0.1, 0.2, size=s)),
        ("Unimodality", lambda: P.Unimodality(), unimodal, lambda s: np.tile([[1.0], [0.0]], (s[0] // 2 + 1, s[1]))[:s[0]]),
        ("Unimodality_nn", lambda: P.Unimodality(non_negativity=True), unimodal, lambda s: -unimodal(s) - 0.1),
        ("UnitSimplex", lambda: P.UnitSimplex(), simplex, lambda s: rs.uniform(1, 2, size=s)),
        ("GeneralizedL2", lambda: P.GeneralizedL2Penalty(_lap(J)), lambda s: np.ones(s) * rs.uniform(size=(1, s[1])),
         lambda s: rs.standard_normal(size=s)),
        ("TotalVariation", lambda: P.TotalVariationPenalty(0.5), lambda s: np.ones(s) * rs.uniform(size=(1, s[1])),
         lambda s: rs.standard_normal(size=s) * 3),
        ("TotalVariation_l1", lambda: P.TotalVariationPenalty(0.5, l1_strength=0.3), lambda s: np.zeros(s),
         lambda s: rs.standard_normal(size=s) * 3),
    ], (J, R)


CASES, (J, R) = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_penalty_contract(case):
    from matcouply_b200 import penalties as P

    name, make, invariant, non_invariant = case
    pen = make()
    rs = np.random.RandomState(1)
    shapes = [(J, R)] * 4
    # invariant points stay (matrix, list of matrices, and rows where the class has a row update)
    inv = [invariant(s) for s in shapes]
    out = pen.factor_matrices_update(inv, [10] * 4, [None] * 4)
    for a, b in zip(inv, out):
        np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(pen.factor_matrix_update(inv[0], 10, None), inv[0], rtol=1e-9, atol=1e-9)
    # non-invariant points move
    non = [non_invariant(s) for s in shapes]
    out = pen.factor_matrices_update(non, [10] * 4, [None] * 4)
    for a, b in zip(non, out):
        assert not np.allclose(b, a, rtol=1e-6, atol=1e-6)
    if isinstance(pen, P.RowVectorPenalty):
        np.testing.assert_allclose(pen.factor_matrix_row_update(inv[0][0], 10, None), inv[0][0], rtol=1e-9, atol=1e-9)
        assert not np.allclose(pen.factor_matrix_row_update(non[0][0], 10, None), non[0][0])
    # the prox does not increase the penalty; hard constraints report zero
    rnd = [rs.standard_normal(size=s) for s in shapes]
    out = pen.factor_matrices_update(rnd, [10] * 4, [None] * 4)
    assert pen.penalty(out) <= pen.penalty(rnd) + 1e-12
    if isinstance(pen, P.HardConstraintMixin):
        assert pen.penalty(rnd) == 0 and pen.penalty(rnd[0]) == 0
    # shifting and aux conversion (penalties.py:268-343)
    aux, dual = rs.standard_normal(size=(J, R)), rs.standard_normal(size=(J, R))
    np.testing.assert_allclose(pen.subtract_from_aux(aux, dual), aux - dual)
    assert pen.aux_as_matrix(aux) is aux
    # aux / dual initialisation (penalties.py:36-261): shapes per mode, given arrays, errors
    mats = [rs.standard_normal(size=(J, 5)) for _ in range(4)]
    for init in ("random_uniform", "random_standard_normal", "zeros"):
        p2 = make()
        p2.aux_init = p2.dual_init = init
        assert p2.init_aux(mats, R, 0, random_state=1).shape == (4, R)
        assert p2.init_dual(mats, R, 2, random_state=1).shape == (5, R)
        lst = p2.init_aux(mats, R, 1, random_state=1)
        assert len(lst) == 4 and all(a.shape == (J, R) for a in lst)
        if init == "zeros":
            assert not np.any(lst[0])
    p3 = make()
    p3.aux_init = np.ones((4, R))
    np.testing.assert_array_equal(p3.init_aux(mats, R, 0), np.ones((4, R)))
    p3.aux_init = np.ones((4, R + 1))
    with pytest.raises(ValueError):
        p3.init_aux(mats, R, 0)
    p3.aux_init = "no such init"
    with pytest.raises(ValueError):
        p3.init_aux(mats, R, 0)
    with pytest.raises(TypeError):
        make().init_aux(mats, 3.0, 0)
    with pytest.raises(ValueError):
        make().init_aux(mats, R, 3)


def test_parafac2_contract():
    """tests/test_penalties.py:432-519: P_i Delta is a fixed point, more iterations do not fit worse, type errors."""
    from matcouply_b200 import penalties as P

    rs = np.random.RandomState(2)
    Rk, Js = 3, [5, 8, 6, 9]
    delta = rs.standard_normal(size=(Rk, Rk))
    bases = [np.linalg.qr(rs.standard_normal(size=(j, Rk)))[0] for j in Js]
    inv = [b @ delta for b in bases]
    pen = P.Parafac2()
    out_b, out_d = pen.factor_matrices_update(inv, [10] * 4, (bases, delta))
    for a, b in zip(inv, pen.auxes_as_matrices((out_b, out_d))):
        np.testing.assert_allclose(b, a, atol=1e-9)
    rnd = [rs.standard_normal(size=(j, Rk)) for j in Js]
    err = {}
    for n in (1, 5):
        ob, od = P.Parafac2(n_iter=n).factor_matrices_update(rnd, [10] * 4, (bases, delta))
        err[n] = sum(np.sum((m - z) ** 2) for m, z in zip(rnd, pen.auxes_as_matrices((ob, od))))
    assert err[5] <= err[1] + 1e-12
    assert pen.penalty(rnd) == 0
    with pytest.raises(TypeError):
        pen.penalty(rnd[0])
    with pytest.raises(TypeError):
        pen.subtract_from_aux(rnd[0], rnd[0])
    with pytest.raises(TypeError):
        pen.aux_as_matrix(rnd[0])
    with pytest.raises(ValueError):
        pen.init_aux(rnd, Rk, 0)
    b0, d0 = pen.init_aux([np.zeros((j, 7)) for j in Js], Rk, 1, random_state=0)
    assert d0.shape == (Rk, Rk) and all(np.array_equal(np.asarray(b), np.eye(j, Rk)) for b, j in zip(b0, Js))
