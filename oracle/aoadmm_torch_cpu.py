"""ORACLE (torch-CPU flavour) — the reference AO-ADMM path as it runs under TensorLy's PyTorch backend on the host.

TEST INFRASTRUCTURE ONLY, like ``aoadmm_oracle.py``: used by ``tests/`` (checked against the NumPy oracle) and by
``bench.py``'s ``cpu_baseline`` legs as the "torch-CPU" column BASELINE.json's north_star asks for next to the NumPy
one.  Nothing under ``matcouply_b200/`` imports it.  It never touches a CUDA device.

The reference has ONE source for both backends (/root/reference/src/matcouply/decomposition.py, penalties.py); under
the PyTorch backend every ``tl.*`` call becomes the ``torch.*`` call of the same name (``tl.dot`` ->
``torch.matmul``, ``tl.clip`` -> ``torch.clamp``, ``truncated_svd`` -> ``torch.linalg.svd`` truncated, ...).  This
module restates that run for the penalties the torch backend supports on the BASELINE configs — NonNegativity
(penalties.py:488-508), L1Penalty (:545-592) and Parafac2 (:1018-1324); Unimodality refuses any non-NumPy backend
(:1008-1009) and TV / UnitSimplex call NumPy-only code, so configs 0 and 3 have no torch-CPU column.

Sub-solvers follow decomposition.py:222-292 (B), :295-344 (C), :120-219 (A) line by line, the diagnostics :351-452;
the initial state is drawn by the NumPy oracle's own code (the reference draws it from ``np.random.RandomState`` under
every backend, decomposition.py:31-39) and converted, so fp64 trajectories agree with the NumPy oracle to round-off
(``tests/test_oracle.py::test_torch_cpu_oracle_matches_numpy_oracle``).
"""
import numpy as np
import torch

from . import aoadmm_oracle as O


def _svd(M):
    return torch.linalg.svd(M, full_matrices=False)


class _NonNeg:
    def __init__(self, src):
        pass

    def shifted(self, aux, dual):
        return aux - dual

    def shifted_list(self, auxes, duals):
        return [a - d for a, d in zip(auxes, duals)]

    def prox(self, M, rho, aux):
        return torch.clamp(M, 0, float("inf"))

    prox_row = prox

    def prox_list(self, Ms, rhos, auxes):
        return [self.prox(M, rho, aux) for M, rho, aux in zip(Ms, rhos, auxes)]

    def value(self, x):
        return 0


class _L1(_NonNeg):
    def __init__(self, src):
        self.strength, self.non_negativity = src.strength, src.non_negativity

    def prox(self, M, rho, aux):  # penalties.py:570-587
        if self.non_negativity:
            return torch.clamp(M - self.strength / rho, 0, float("inf"))
        return torch.sign(M) * torch.clamp(torch.abs(M) - self.strength / rho, 0, float("inf"))

    prox_row = prox

    def value(self, x):  # penalties.py:589-592
        if isinstance(x, list):
            return float(sum(torch.sum(torch.abs(xi)) for xi in x) * self.strength)
        return float(torch.sum(torch.abs(x)) * self.strength)


class _Parafac2(_NonNeg):
    def shifted_list(self, auxes, duals):  # penalties.py:1256-1281
        P, delta = auxes
        return [torch.matmul(Pi, delta) - d for Pi, d in zip(P, duals)]

    def prox_list(self, Ms, rhos, auxes):  # penalties.py:1224-1250 (n_iter = 1)
        _, delta = auxes
        R = delta.shape[0]
        P = []
        for M in Ms:
            U, _, Vh = _svd(torch.matmul(M, delta.T))
            P.append(torch.matmul(U[:, :R], Vh[:R, :]))
        new_delta = 0
        for M, Pi, rho in zip(Ms, P, rhos):
            new_delta = new_delta + rho * torch.matmul(Pi.T, M)
        return P, new_delta / sum(rhos)


_TORCH_CLASSES = {O.NonNeg: _NonNeg, O.L1P: _L1, O.Parafac2P: _Parafac2}


def _solve_B(X, regs, A, Bs, C, auxes, duals, n_inner):  # decomposition.py:222-292
    R = A.shape[1]
    eye = torch.eye(R, dtype=A.dtype)
    CtC = torch.matmul(C.T, C)
    rhs = [torch.matmul(Xi, C * a) for Xi, a in zip(X, A)]
    lhs = [((CtC * a).T * a).T for a in A]
    rhos = [0.5 * torch.trace(L) for L in lhs]
    svds = [_svd(L + eye * (rho * len(regs))) for L, rho in zip(lhs, rhos)]
    Bs = list(Bs)
    for _ in range(n_inner):
        shifted = [reg.shifted_list(aux, dual) for reg, aux, dual in zip(regs, auxes, duals)]
        for i in range(len(X)):
            U, s, Uh = svds[i]
            acc = 0
            for sh in shifted:
                acc = acc + sh[i]
            Bs[i] = torch.matmul(torch.matmul(rhos[i] * acc + rhs[i], U / s), Uh)
        for n, reg in enumerate(regs):
            auxes[n] = reg.prox_list([B + d for B, d in zip(Bs, duals[n])], rhos, auxes[n])
            sh = reg.shifted_list(auxes[n], duals[n])
            duals[n] = [B - s_ for B, s_ in zip(Bs, sh)]
    return Bs, auxes, duals


def _solve_C(X, regs, A, Bs, C, auxes, duals, n_inner):  # decomposition.py:295-344
    R = C.shape[1]
    lhs, rhs = 0, 0
    for Xi, B, a in zip(X, Bs, A):
        Ba = B * a
        lhs = lhs + torch.matmul(Ba.T, Ba)
        rhs = rhs + torch.matmul(Xi.T, Ba)
    rho = 0.5 * torch.trace(lhs)
    U, s, Uh = _svd(lhs + torch.eye(R, dtype=C.dtype) * (rho * len(regs)))
    for _ in range(n_inner):
        acc = 0
        for reg, aux, dual in zip(regs, auxes, duals):
            acc = acc + reg.shifted(aux, dual)
        C = torch.matmul(torch.matmul(acc * rho + rhs, U / s), Uh)
        for n, reg in enumerate(regs):
            auxes[n] = reg.prox(C + duals[n], rho, auxes[n])
            duals[n] = C - reg.shifted(auxes[n], duals[n])
    return C, auxes, duals


def _solve_A(X, regs, A, Bs, C, auxes, duals, n_inner):  # decomposition.py:120-219
    R, K = A.shape[1], C.shape[0]
    eye = torch.eye(R, dtype=A.dtype)
    CtC = torch.matmul(C.T, C)
    cross, rhs = [], []
    for Xi, B in zip(X, Bs):
        if B.shape[0] > K:
            BtXC = torch.matmul(torch.matmul(B.T, Xi), C)
        else:
            BtXC = torch.matmul(B.T, torch.matmul(Xi, C))
        cross.append(torch.matmul(B.T, B) * CtC)
        rhs.append(torch.diag(BtXC))
    rhos = [0.5 * torch.trace(L) for L in cross]
    svds = [_svd(L + eye * (rho * len(regs))) for L, rho in zip(cross, rhos)]
    A = A.clone()
    for _ in range(n_inner):
        shifted = [reg.shifted(aux, dual) for reg, aux, dual in zip(regs, auxes, duals)]
        for i in range(len(X)):
            U, s, Uh = svds[i]
            acc = 0
            for sh in shifted:
                acc = acc + sh[i]
            A[i, :] = torch.matmul(torch.matmul(rhos[i] * acc + rhs[i], U / s), Uh)
        for n, reg in enumerate(regs):
            moved = A + duals[n]
            new_aux = auxes[n].clone()
            for i, rho in enumerate(rhos):
                new_aux[i, :] = reg.prox_row(moved[i], rho, auxes[n][i])
            auxes[n] = new_aux
            duals[n] = A - reg.shifted(auxes[n], duals[n])
    return A, auxes, duals, (rhs, cross)


def _rss(xs):
    return torch.sqrt(sum(torch.sum(x ** 2) for x in xs))


def ao_admm_torch_cpu(matrices, rank, n_iter_max=1, non_negative=None, parafac2=None, l1_penalty=None,
                      random_state=0, dtype=torch.float64, inner_n_iter_max=5):
    """``n_iter_max`` outer iterations with the default options of ``cmf_aoadmm`` (no stopping rule: the bench and the
    test run a fixed count).  Returns dict(A, B_is, C, rec_errors, regularized_loss) as NumPy / floats."""
    rs = np.random.RandomState(random_state)
    mats = [np.asarray(M, dtype=np.float64) for M in matrices]
    I, K = len(mats), mats[0].shape[1]
    A = rs.uniform(size=(I, rank))
    C = rs.uniform(size=(K, rank))
    Bs = [rs.uniform(size=(M.shape[0], rank)) for M in mats]
    src = O.build_penalties(non_negative=non_negative, parafac2=parafac2, l1_penalty=l1_penalty)
    for mode in src:
        for r in mode:
            if type(r) not in _TORCH_CLASSES:
                raise ValueError(f"{type(r).__name__} has no torch-backend path in the reference")
    aux = [[r.init_aux(mats, rank, m, rs) for r in src[m]] for m in range(3)]
    dual = [[r.init_dual(mats, rank, m, rs) for r in src[m]] for m in range(3)]

    def t(v):
        if isinstance(v, np.ndarray):
            return torch.as_tensor(v, dtype=dtype)
        if isinstance(v, (list, tuple)):
            return type(v)(t(u) for u in v)
        return v

    regs = [[_TORCH_CLASSES[type(r)](r) for r in mode] for mode in src]
    X, A, C, Bs, aux, dual = t(mats), t(A), t(C), t(Bs), t(aux), t(dual)
    aux, dual = [list(a) for a in aux], [list(d) for d in dual]
    norm_X = _rss(X)
    rec_errors, losses = [], []
    for _ in range(n_iter_max):
        Bs, aux[1], dual[1] = _solve_B(X, regs[1], A, Bs, C, aux[1], dual[1], inner_n_iter_max)
        C, aux[2], dual[2] = _solve_C(X, regs[2], A, Bs, C, aux[2], dual[2], inner_n_iter_max)
        A, aux[0], dual[0], (rhs, cross) = _solve_A(X, regs[0], A, Bs, C, aux[0], dual[0], inner_n_iter_max)
        # decomposition.py:351-417: the gaps are evaluated every iteration (their cost belongs to the iteration)
        for m, x in ((0, A), (2, C)):
            for reg, z in zip(regs[m], aux[m]):
                torch.sqrt(torch.sum(reg.shifted(z, x) ** 2)) / torch.sqrt(torch.sum(x ** 2))
        for reg, z in zip(regs[1], aux[1]):
            _rss(reg.shifted_list(z, Bs)) / _rss(Bs)
        # decomposition.py:445-452 (fit term from the A-update's intermediates) and :1016-1023 (loss)
        inner = sum(torch.sum(r * a) for r, a in zip(rhs, A))
        ncmf = sum(torch.sum(torch.diag(a) @ cross[i] @ torch.diag(a)) for i, a in enumerate(A))
        err = float(torch.sqrt(torch.clamp(norm_X ** 2 - 2 * inner + ncmf, min=0)) / norm_X)
        rec_errors.append(err)
        losses.append(0.5 * err ** 2 + sum(r.value(A) for r in regs[0]) + sum(r.value(Bs) for r in regs[1])
                      + sum(r.value(C) for r in regs[2]))
    return dict(A=A.numpy(), B_is=[B.numpy() for B in Bs], C=C.numpy(), rec_errors=rec_errors,
                regularized_loss=losses)
