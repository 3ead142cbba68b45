"""ORACLE — CPU (NumPy float64) restatement of the reference AO-ADMM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``matcouply_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and
there only as the checker / the CPU baseline, never as the product path.

What it restates (paths relative to /root/reference/src/matcouply):

* outer AO-ADMM driver and stopping rule ......... decomposition.py:862-1100  -> :func:`ao_admm`
* B-, C-, A-mode ADMM sub-solvers ................ decomposition.py:222-292, 295-344, 120-219
* feasibility gaps / fit term / loss ............. decomposition.py:351-417, 420-452, 617-627, 1016-1023
* keyword -> penalty parsing ..................... decomposition.py:455-467, 470-614
* random init of factors / aux / dual ............ decomposition.py:31-39, 78-89; penalties.py:125-147, 239-261, 1161-1175
* prox operators ................................. penalties.py:488-508 (NonNegativity), 511-542 (Box), 545-592 (L1),
                                                   844-925 (L2Ball), 983-1015 (Unimodality), 1018-1324 (Parafac2)
* unimodal regression ............................ _unimodal_regression.py:24-141 (C restatement in
                                                   oracle/unimodal_oracle.c; pure-Python twin below for tiny cases)

The arithmetic follows the reference operation by operation (same ``np.dot`` calls, same LAPACK SVD based
solve ``x (U/s) Uh``, same summation order over slices), so on identical inputs and ``random_state`` it
reproduces the reference to round-off.  PINNED: ``tests/test_oracle.py`` checks it against golden trajectories
generated from the unmodified reference (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``), and, when
``/root/reference`` is present, against the reference run live.

Symbols follow the reference code: ``I`` slices ``X_i`` of shape ``J_i x K``; ``A`` is ``I x R``;
``B_i`` is ``J_i x R``; ``C`` is ``K x R``; modes 0/1/2 = A/B/C.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _load_c():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.oracle_unimodal_regression.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            lib.oracle_unimodal_regression.restype = None
            _LIB = lib
        else:
            _LIB = False
    return _LIB


# ----------------------------------------------------------------------------------------------------------
# unimodal regression (_unimodal_regression.py)
# ----------------------------------------------------------------------------------------------------------
def _prefix_isotonic_py(y, non_negativity):
    """Pure-Python twin of ``prefix_isotonic_regression`` (_unimodal_regression.py:24-69); tiny inputs only."""
    n = y.shape[0]
    swy = y.astype(np.float64).copy()
    swy2 = swy * y
    sw = np.ones(n)
    level = np.zeros(n)
    start = np.zeros(n, dtype=np.int64)
    err = np.zeros(n + 1)
    level[0] = y[0]
    neg = np.zeros(n, dtype=bool)
    if non_negativity:
        csum = np.cumsum(swy2)
        if level[0] < 0:
            neg[0] = True
            err[1] = csum[0]
    for i in range(1, n):
        level[i] = y[i]
        start[i] = i
        while start[i] != 0 and level[i] <= level[start[i] - 1]:
            p = start[i] - 1
            swy[i] += swy[p]
            swy2[i] += swy2[p]
            sw[i] += sw[p]
            level[i] = swy[i] / sw[i]
            start[i] = start[p]
        block_err = swy2[i] - (swy[i] ** 2 / sw[i])
        if non_negativity and level[i] < 0:
            neg[i] = True
            err[i + 1] = csum[i]
        else:
            err[i + 1] = block_err + err[start[i]]
    level[neg] = 0.0
    return level, start, err


def _expand_blocks(end, level, start):
    out = np.empty(end)
    idx = end - 1
    while idx >= 0:
        out[start[idx]: idx + 1] = level[idx]
        idx = start[idx] - 1
    return out


def unimodal_column_py(y, non_negativity=False):
    """_unimodal_regression.py:84-104 for one vector. Returns (fit, peak_index, error)."""
    n = y.shape[0]
    lvL, stL, eL = _prefix_isotonic_py(y, non_negativity)
    lvR, stR, eR = _prefix_isotonic_py(y[::-1], non_negativity)
    best, t = eR[-1], 0
    for i in range(n + 1):
        e = eL[i] + eR[n - i]
        if e < best:
            best, t = e, i
    left = _expand_blocks(t, lvL, stL)
    right = _expand_blocks(n - t, lvR, stR)
    return np.concatenate([left, right[::-1]]), t, best


def unimodal_regression(Y, non_negativity=False, return_peaks=False, force_python=False):
    """Column-wise unimodal regression of a (J x R) matrix (_unimodal_regression.py:107-141)."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    one_d = Y.ndim == 1
    M = Y.reshape(Y.shape[0], -1)
    out = np.empty_like(M)
    peaks = np.zeros(M.shape[1], dtype=np.int32)
    lib = None if force_python else _load_c()
    if lib:
        errs = np.zeros(M.shape[1])
        lib.oracle_unimodal_regression(M.ctypes.data, M.shape[0], M.shape[1], int(bool(non_negativity)),
                                       out.ctypes.data, peaks.ctypes.data, errs.ctypes.data)
    else:
        for r in range(M.shape[1]):
            out[:, r], peaks[r], _ = unimodal_column_py(M[:, r].copy(), non_negativity)
    out = out.reshape(Y.shape) if not one_d else out[:, 0]
    return (out, peaks) if return_peaks else out


# ----------------------------------------------------------------------------------------------------------
# penalties (penalties.py) — only the behaviour the AO-ADMM loop touches
# ----------------------------------------------------------------------------------------------------------
def _draw(rs, how, shape):
    if how == "random_uniform":
        return rs.uniform(size=shape)
    if how == "random_standard_normal":
        return rs.standard_normal(size=shape)
    if how == "zeros":
        return np.zeros(shape)
    raise ValueError("Unknown aux init: {}".format(how))


class OraclePenalty:
    """Common protocol (penalties.py:21-366).  ``aux_init``/``dual_init`` are strings or given arrays."""

    name = "penalty"
    matrixwise = False  # True => no row update available (MatrixPenalty subclasses: L2Ball, Unimodality)

    def __init__(self, aux_init="random_uniform", dual_init="random_uniform"):
        self.aux_init, self.dual_init = aux_init, dual_init

    # penalties.py:36-147 / 149-261 (shape validation is host logic of the product; the oracle trusts inputs)
    def _init(self, how, matrices, rank, mode, rs):
        if not isinstance(how, str):
            return how
        if mode == 0:
            return _draw(rs, how, (len(matrices), rank))
        if mode == 2:
            return _draw(rs, how, (matrices[0].shape[1], rank))
        return [_draw(rs, how, (M.shape[0], rank)) for M in matrices]

    def init_aux(self, matrices, rank, mode, rs):
        return self._init(self.aux_init, matrices, rank, mode, rs)

    def init_dual(self, matrices, rank, mode, rs):
        return self._init(self.dual_init, matrices, rank, mode, rs)

    def shifted(self, aux, dual):  # subtract_from_aux, penalties.py:290-309
        return aux - dual

    def shifted_list(self, auxes, duals):  # subtract_from_auxes, penalties.py:268-288
        return [self.shifted(a, d) for a, d in zip(auxes, duals)]

    def value(self, x):  # penalty(); hard constraints are 0 (penalties.py:469-485)
        return 0

    def prox(self, M, rho, aux):  # factor_matrix_update
        raise NotImplementedError

    def prox_row(self, row, rho, aux_row):  # factor_matrix_row_update
        return self.prox(row, rho, aux_row)

    def prox_list(self, Ms, rhos, auxes):  # factor_matrices_update, penalties.py:392-408
        return [self.prox(M, rho, a) for M, rho, a in zip(Ms, rhos, auxes)]


class NonNeg(OraclePenalty):  # penalties.py:488-508
    name = "nonneg"

    def prox(self, M, rho, aux):
        return np.clip(M, 0, float("inf"))


class BoxP(OraclePenalty):  # penalties.py:511-542
    name = "box"

    def __init__(self, lo, hi, **kw):
        super().__init__(**kw)
        self.lo, self.hi = lo, hi

    def prox(self, M, rho, aux):
        return np.clip(M, self.lo, self.hi)


class L1P(OraclePenalty):  # penalties.py:545-592
    name = "l1"

    def __init__(self, strength, non_negativity=False, **kw):
        super().__init__(**kw)
        if strength < 0:
            raise ValueError("Regularization strength must be nonnegative.")
        self.strength, self.non_negativity = strength, non_negativity

    def prox(self, M, rho, aux):
        if self.non_negativity:
            return np.clip(M - self.strength / rho, 0, float("inf"))
        return np.sign(M) * np.clip(np.abs(M) - self.strength / rho, 0, float("inf"))

    def value(self, x):
        if isinstance(x, np.ndarray):
            return np.sum(np.abs(x)) * self.strength
        return sum(np.sum(np.abs(xi)) for xi in x) * self.strength


class L2BallP(OraclePenalty):  # penalties.py:844-925
    name = "l2ball"
    matrixwise = True

    def __init__(self, bound, non_negativity=False, **kw):
        super().__init__(**kw)
        if bound <= 0:
            raise ValueError("The norm bound must be positive.")
        self.bound, self.non_negativity = bound, non_negativity

    def prox(self, M, rho, aux):
        if self.non_negativity:
            M = np.clip(M, 0, float("inf"))
        norms = np.sqrt(np.sum(M ** 2, axis=0))
        norms = np.clip(norms, self.bound, float("inf"))
        return M * self.bound / norms

    def prox_row(self, row, rho, aux_row):
        raise AttributeError("L2Ball has no row update (needs constant_feasibility_penalty on mode 0)")


class UnimodalP(OraclePenalty):  # penalties.py:983-1015
    name = "unimodal"
    matrixwise = True

    def __init__(self, non_negativity=False, **kw):
        super().__init__(**kw)
        self.non_negativity = non_negativity

    def prox(self, M, rho, aux):
        return unimodal_regression(M, non_negativity=self.non_negativity)

    def prox_row(self, row, rho, aux_row):
        raise AttributeError("Unimodality has no row update")


class Parafac2P(OraclePenalty):  # penalties.py:1018-1324
    name = "parafac2"
    matrixwise = True

    def __init__(self, n_iter=1, update_basis_matrices=True, update_coordinate_matrix=True, **kw):
        super().__init__(**kw)
        self.n_iter = n_iter
        self.update_basis_matrices, self.update_coordinate_matrix = update_basis_matrices, update_coordinate_matrix

    def init_aux(self, matrices, rank, mode, rs):  # :1161-1175, tuple init :1176-1220
        if isinstance(self.aux_init, tuple):
            return self.aux_init
        delta = _draw(rs, self.aux_init, (rank, rank))
        return [np.eye(M.shape[0], rank) for M in matrices], delta

    def shifted_list(self, auxes, duals):  # :1256-1281
        P, delta = auxes
        return [np.dot(Pi, delta) - d for Pi, d in zip(P, duals)]

    def prox_list(self, Ms, rhos, auxes):  # :1224-1250
        P, delta = auxes
        R = delta.shape[0]
        for _ in range(self.n_iter):
            if self.update_basis_matrices:
                P = []
                for M in Ms:
                    T = np.matmul(M, delta.T)
                    U, _, Vh = np.linalg.svd(T, full_matrices=R > min(T.shape))
                    P.append(np.matmul(U[:, :R], Vh[:R, :]))
            if self.update_coordinate_matrix:
                new_delta = 0
                for M, Pi, rho in zip(Ms, P, rhos):
                    new_delta += rho * Pi.T @ M
                delta = new_delta / sum(rhos)
            if (not self.update_coordinate_matrix) or (not self.update_basis_matrices):
                break
        return P, delta

    def as_matrices(self, auxes):  # auxes_as_matrices :1287-1304
        P, delta = auxes
        return [np.dot(Pi, delta) for Pi in P]


class GeneralizedL2P(OraclePenalty):  # penalties.py:595-747
    """x^T M x on every column; prox = U diag(1 / (s + rho/2)) U^T (rho/2 x) with the SVD of M (:724-730)."""
    name = "generalized_l2"
    matrixwise = True

    def __init__(self, norm_matrix, **kw):
        super().__init__(**kw)
        self.norm_matrix = np.asarray(norm_matrix, dtype=np.float64)
        if not np.all(self.norm_matrix.T == self.norm_matrix) or np.any(np.linalg.eigvals(self.norm_matrix) < -1e-14):
            raise ValueError("The norm matrix should be symmetric positive semidefinite")
        self._U, self._s, _ = np.linalg.svd(self.norm_matrix, full_matrices=False)

    def prox(self, M, rho, aux):
        s_aug = self._s + 0.5 * rho
        tmp = 0.5 * rho * M
        tmp = np.dot(self._U.T, tmp)
        return np.dot(self._U * (1 / s_aug), tmp)

    def prox_row(self, row, rho, aux_row):
        raise AttributeError("GeneralizedL2Penalty has no row update")

    def _value(self, x):
        return np.trace(np.dot(np.dot(x.T, self.norm_matrix), x))

    def value(self, x):
        if isinstance(x, np.ndarray):
            return self._value(x)
        return sum(self._value(xi) for xi in x)


def _scipy_bisect(f, xa, xb, xtol=2e-12, rtol=8.881784197001252e-16, maxiter=100):
    """scipy.optimize.bisect's loop (scipy/optimize/Zeros/bisect.c), restated."""
    fa, fb = f(xa), f(xb)
    if fa == 0:
        return xa
    if fb == 0:
        return xb
    if np.signbit(fa) == np.signbit(fb):
        raise ValueError("f(a) and f(b) must have different signs")
    dm = xb - xa
    for _ in range(maxiter):
        dm *= 0.5
        xm = xa + dm
        fm = f(xm)
        if fm * fa >= 0:
            xa = xm
        if fm == 0 or abs(dm) < xtol + rtol * abs(xm):
            return xm
    raise RuntimeError("bisect failed to converge")


class UnitSimplexP(OraclePenalty):  # penalties.py:928-980
    name = "unit_simplex"
    matrixwise = True

    @staticmethod
    def multiplier(col):  # :941-969
        min_val = np.min(col) - 1
        max_val = np.max(col)
        min_val -= 1e-5
        min_val = min(0.9 * min_val, 1.1 * min_val)
        max_val += 1e-5
        max_val = max(0.9 * max_val, 1.1 * max_val)
        return _scipy_bisect(lambda mu: np.sum(np.clip(col - mu, 0, None)) - 1, min_val, max_val)

    def prox(self, M, rho, aux):  # :971-980
        out = np.zeros(M.shape)
        for r in range(M.shape[1]):
            out[:, r] = np.clip(M[:, r] - self.multiplier(M[:, r]), 0, None)
        return out

    def prox_row(self, row, rho, aux_row):
        raise AttributeError("UnitSimplex has no row update")


def tv_denoise_1d(y, lam):
    """argmin_x 0.5 ||x - y||^2 + lam sum_k |x[k+1] - x[k]|: Condat's direct algorithm (IEEE SPL 20(11), 2013), the
    algorithm of the reference's un-vendored `condat_tv` dependency (penalties.py:6-11, 821-823).  PARITY UNPINNED
    against that package (not installable offline); pinned instead by the KKT certificate of the unique minimiser
    (tests/test_oracle.py::test_tv_oracle_kkt)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    x = np.zeros(n)
    if n == 0:
        return x
    k = k0 = kplus = kminus = 0
    umin, umax = lam, -lam
    vmin, vmax = y[0] - lam, y[0] + lam
    while True:
        done = False
        while k == n - 1:
            if umin < 0.0:
                while True:
                    x[k0] = vmin
                    k0 += 1
                    if k0 > kminus:
                        break
                k = kminus = k0
                vmin = y[k0]
                umin = lam
                umax = vmin + umin - vmax
            elif umax > 0.0:
                while True:
                    x[k0] = vmax
                    k0 += 1
                    if k0 > kplus:
                        break
                k = kplus = k0
                vmax = y[k0]
                umax = -lam
                umin = vmax + umax - vmin
            else:
                vmin += umin / (k - k0 + 1)
                x[k0:k + 1] = vmin
                done = True
                break
        if done:
            return x
        umin += y[k + 1] - vmin
        if umin < -lam:
            x[k0:kminus + 1] = vmin
            k0 = kminus + 1
            k = kplus = kminus = k0
            vmin = y[k0]
            vmax = vmin + 2.0 * lam
            umin, umax = lam, -lam
            continue
        umax += y[k + 1] - vmax
        if umax > lam:
            x[k0:kplus + 1] = vmax
            k0 = kplus + 1
            k = kplus = kminus = k0
            vmax = y[k0]
            vmin = vmax - 2.0 * lam
            umin, umax = lam, -lam
            continue
        k += 1
        if umin >= lam:
            kminus = k
            vmin += (umin - lam) / (kminus - k0 + 1)
            umin = lam
        if umax <= -lam:
            kplus = k
            vmax += (umax + lam) / (kplus - k0 + 1)
            umax = -lam


class TotalVariationP(OraclePenalty):  # penalties.py:750-841
    name = "total_variation"
    matrixwise = True

    def __init__(self, reg_strength, l1_strength=0, **kw):
        super().__init__(**kw)
        if reg_strength <= 0:
            raise ValueError("The TV regularization strength must be positive.")
        if l1_strength < 0:
            raise ValueError("The L1 regularization strength must be non-negative.")
        self.reg_strength, self.l1_strength = reg_strength, l1_strength

    def prox(self, M, rho, aux):  # :819-827 (note the factor 2 of :822)
        X = np.stack([tv_denoise_1d(M[:, r], self.reg_strength * 2 / rho) for r in range(M.shape[1])], axis=1)
        if self.l1_strength:
            return np.sign(X) * np.clip(np.abs(X) - self.l1_strength / rho, 0, float("inf"))
        return X

    def prox_row(self, row, rho, aux_row):
        raise AttributeError("TotalVariationPenalty has no row update")

    def _value(self, x):  # :829-834
        v = self.reg_strength * np.sum(np.abs(np.diff(x, axis=0)))
        if self.l1_strength:
            v = v + self.l1_strength * np.sum(np.abs(x))
        return v

    def value(self, x):
        if isinstance(x, np.ndarray):
            return self._value(x)
        return sum(self._value(xi) for xi in x)


ORACLE_CLASSES = {"NonNegativity": NonNeg, "Box": BoxP, "L1Penalty": L1P, "L2Ball": L2BallP, "Unimodality": UnimodalP,
                  "Parafac2": Parafac2P, "GeneralizedL2Penalty": GeneralizedL2P, "UnitSimplex": UnitSimplexP,
                  "TotalVariationPenalty": TotalVariationP}


def regs_from_spec(spec, classes=None):
    """[[["ClassName", {kwargs}], ...] per mode] -> penalty objects (JSON-safe description of a `regs` argument).
    ``classes``: name -> class mapping (default: the oracle's); array-valued kwargs arrive as nested lists."""
    oracle_side = classes is None
    classes = ORACLE_CLASSES if classes is None else classes
    # the spec uses the reference's constructor keywords; the oracle classes name some of them differently
    rename = {"L2Ball": {"norm_bound": "bound"}, "Box": {"min_val": "lo", "max_val": "hi"},
              "L1Penalty": {"reg_strength": "strength"}}
    out = []
    for mode_spec in spec:
        mode = []
        for name, kw in mode_spec:
            kw = {k: (np.asarray(v, dtype=np.float64) if isinstance(v, list) else v) for k, v in kw.items()}
            if oracle_side:
                kw = {rename.get(name, {}).get(k, k): v for k, v in kw.items()}
            cls = classes[name] if isinstance(classes, dict) else getattr(classes, name)
            mode.append(cls(**kw))
        out.append(mode)
    return out


def _listify(v, name):  # decomposition.py:455-467
    if hasattr(v, "get"):
        return [v.get(i, None) for i in range(3)]
    try:
        iter(v)
    except TypeError:
        return [v] * 3
    out = list(v)
    if len(out) != 3:
        raise ValueError(
            "All parameters must be a dictionary, non-iterable value or non-dictionary iterable of length 3."
            f" {name} is iterable of length {len(out)}."
        )
    return out


def build_penalties(non_negative=None, lower_bound=None, upper_bound=None, l2_norm_bound=None, unimodal=None,
                    parafac2=None, l1_penalty=None, aux_init="random_uniform", dual_init="random_uniform",
                    generalized_l2_penalty=None, tv_penalty=None):
    """decomposition.py:470-614 restricted to the in-scope penalties; fixed per-mode order
    Parafac2, Unimodality, L2Ball, L1, Box, NonNegativity; ``non_negative`` folded into the others."""
    nn = _listify(non_negative, "non_negative")
    ub = _listify(upper_bound, "upper_bound")
    lb = _listify(lower_bound, "lower_bound")
    l2b = _listify(l2_norm_bound, "l2_norm_bound")
    uni = _listify(unimodal, "unimodal")
    pf2 = [False, bool(parafac2), False]
    l1 = _listify(l1_penalty, "l1_penalty")
    gl2 = _listify(generalized_l2_penalty, "generalized_l2_penalty")
    tv = _listify(tv_penalty, "tv_penalty")
    kw = dict(aux_init=aux_init, dual_init=dual_init)
    regs = []
    for m in range(3):
        mode_regs, skip_nn = [], False
        l1m = l1[m] if l1[m] else 0
        if pf2[m]:
            mode_regs.append(Parafac2P(**kw))
        if uni[m]:
            mode_regs.append(UnimodalP(non_negativity=nn[m], **kw))
            skip_nn = True
        if gl2[m] is not None and gl2[m] is not False:  # decomposition.py:583-586
            mode_regs.append(GeneralizedL2P(gl2[m], **kw))
        if l2b[m]:
            mode_regs.append(L2BallP(l2b[m], non_negativity=nn[m], **kw))
            skip_nn = True
        if tv[m]:  # decomposition.py:591-595: the L1 strength moves into the TV penalty
            mode_regs.append(TotalVariationP(tv[m], l1_strength=l1m, **kw))
            l1m = 0
        if l1m:
            mode_regs.append(L1P(l1m, non_negativity=nn[m], **kw))
            skip_nn = True
        if lb[m] is not None or ub[m] is not None:
            lo = -float("inf") if lb[m] is None else lb[m]
            if nn[m]:
                lo = max(lo, 0)
            mode_regs.append(BoxP(lo, ub[m], **kw))
            skip_nn = True
        if nn[m] and not skip_nn:
            mode_regs.append(NonNeg(**kw))
        regs.append(mode_regs)
    return regs


# ----------------------------------------------------------------------------------------------------------
# sub-solvers (decomposition.py:120-344)
# ----------------------------------------------------------------------------------------------------------
def _svd(M):
    return np.linalg.svd(M, full_matrices=False)


def _inner_converged(new, old, regs, auxes, mode, inner_tol):
    """_check_inner_convergence (decomposition.py:92-116)."""
    if not inner_tol or inner_tol < 0:
        return False
    if mode == 1:
        norm = _rss(new)
        change = _rss([b - pb for b, pb in zip(new, old)])
        gaps = [_rss(reg.shifted_list(aux, new)) / norm for reg, aux in zip(regs, auxes)] if regs else []
    else:
        norm = _fro(new)
        change = _fro(new - old)
        gaps = [_fro(reg.shifted(aux, new)) / norm for reg, aux in zip(regs, auxes)] if regs else []
    if change > inner_tol * norm:
        return False
    if len(regs) == 0:
        return True
    return max(gaps) < inner_tol


def solve_mode_B(matrices, regs, A, Bs, C, auxes, duals, l2, n_inner, scale, constant_rho, inner_tol=None):
    """decomposition.py:222-292."""
    R = A.shape[1]
    CtC = np.dot(C.T, C)
    rhs, lhs = [], []
    for X, a in zip(matrices, A):
        rhs.append(np.dot(X, C * a))
        lhs.append(np.transpose(np.transpose(CtC * a) * a))
    rhos = [0.5 * np.trace(L) * scale for L in lhs]
    if constant_rho:
        mx = max(rhos)
        rhos = [mx for _ in rhos]
    lhs = [L + np.eye(R) * (rho * len(regs) + l2) for L, rho in zip(lhs, rhos)]
    svds = [_svd(L) for L in lhs]
    Bs = list(Bs)
    for _ in range(n_inner):
        old_Bs = list(Bs)
        shifted = [reg.shifted_list(aux, dual) for reg, aux, dual in zip(regs, auxes, duals)]
        for i in range(len(matrices)):
            U, s, Uh = svds[i]
            acc = 0
            for sh in shifted:
                acc += sh[i]
            Bs[i] = np.dot(np.dot(rhos[i] * acc + rhs[i], U / s), Uh)
        for n, reg in enumerate(regs):
            moved = [B + d for B, d in zip(Bs, duals[n])]
            auxes[n] = reg.prox_list(moved, rhos, auxes[n])
            sh = reg.shifted_list(auxes[n], duals[n])
            duals[n] = [B - s_ for B, s_ in zip(Bs, sh)]
        if _inner_converged(Bs, old_Bs, regs, auxes, 1, inner_tol):
            break
    return Bs, auxes, duals


def solve_mode_C(matrices, regs, A, Bs, C, auxes, duals, l2, n_inner, scale, inner_tol=None):
    """decomposition.py:295-344."""
    R = C.shape[1]
    lhs, rhs = 0, 0
    for X, B, a in zip(matrices, Bs, A):
        Ba = B * a
        lhs += np.dot(Ba.T, Ba)
        rhs += np.dot(X.T, Ba)
    rho = 0.5 * np.trace(lhs) * scale
    lhs = lhs + np.eye(R) * (rho * len(regs) + l2)
    U, s, Uh = _svd(lhs)
    for _ in range(n_inner):
        old_C = C
        acc = 0
        for reg, aux, dual in zip(regs, auxes, duals):
            acc += reg.shifted(aux, dual)
        C = np.dot(np.dot(acc * rho + rhs, U / s), Uh)
        for n, reg in enumerate(regs):
            auxes[n] = reg.prox(C + duals[n], rho, auxes[n])
            duals[n] = C - reg.shifted(auxes[n], duals[n])
        if _inner_converged(C, old_C, regs, auxes, 2, inner_tol):
            break
    return C, auxes, duals


def solve_mode_A(matrices, regs, A, Bs, C, auxes, duals, l2, n_inner, scale, constant_rho, inner_tol=None):
    """decomposition.py:120-219. Returns also (rhses, cross_products) for the fit term."""
    R = A.shape[1]
    K = C.shape[0]
    CtC = np.dot(C.T, C)
    cross, rhs = [], []
    for X, B in zip(matrices, Bs):
        if B.shape[0] > K:
            BtXC = np.dot(np.dot(B.T, X), C)
        else:
            BtXC = np.dot(B.T, np.dot(X, C))
        cross.append(np.dot(B.T, B) * CtC)
        rhs.append(np.diag(BtXC))
    rhos = [0.5 * np.trace(L) * scale for L in cross]
    if constant_rho:
        mx = max(rhos)
        rhos = [mx for _ in rhos]
    svds = [_svd(L + np.eye(R) * (rho * len(regs) + l2)) for L, rho in zip(cross, rhos)]
    A = A.copy()
    for _ in range(n_inner):
        old_A = A.copy()
        shifted = [reg.shifted(aux, dual) for reg, aux, dual in zip(regs, auxes, duals)]
        for i in range(len(matrices)):
            U, s, Uh = svds[i]
            acc = 0
            for sh in shifted:
                acc += sh[i]
            A[i, :] = np.dot(np.dot(rhos[i] * acc + rhs[i], U / s), Uh)
        for n, reg in enumerate(regs):
            moved = A + duals[n]
            if constant_rho:
                auxes[n] = reg.prox(moved, mx, auxes[n])
            else:
                new_aux = auxes[n].copy()
                for i, rho in enumerate(rhos):
                    new_aux[i, :] = reg.prox_row(moved[i], rho, auxes[n][i])
                auxes[n] = new_aux
            duals[n] = A - reg.shifted(auxes[n], duals[n])
        if _inner_converged(A, old_A, regs, auxes, 0, inner_tol):
            break
    return A, auxes, duals, (rhs, cross)


# ----------------------------------------------------------------------------------------------------------
# diagnostics (decomposition.py:347-452, 617-640)
# ----------------------------------------------------------------------------------------------------------
def _rss(xs):
    return np.sqrt(sum(np.sum(x ** 2) for x in xs))


def _fro(x):
    return np.sqrt(np.sum(np.abs(x) ** 2))


def feasibility_gaps(A, Bs, C, regs, auxA, auxB, auxC):
    """decomposition.py:351-417."""
    An, Bn, Cn = _fro(A), _rss(Bs), _fro(C)
    gA = [_fro(reg.shifted(aux, A)) / An for reg, aux in zip(regs[0], auxA)]
    gB = [_rss(reg.shifted_list(aux, Bs)) / Bn for reg, aux in zip(regs[1], auxB)]
    gC = [_fro(reg.shifted(aux, C)) / Cn for reg, aux in zip(regs[2], auxC)]
    return gA, gB, gC


def reconstruction_error(matrices, A, Bs, C, norm_X, intermediates=None):
    """decomposition.py:420-452 (absolute error, not divided by ||X||)."""
    nx2 = norm_X ** 2
    if intermediates is None:
        ncmf, inner = 0, 0
        CtC = np.dot(C.T, C)
        for i, B in enumerate(Bs):
            Ba = B * A[i]
            if Ba.shape[0] > C.shape[0]:
                inner += np.trace(np.dot(np.dot(Ba.T, matrices[i]), C))
            else:
                inner += np.trace(np.dot(Ba.T, np.dot(matrices[i], C)))
            ncmf += np.sum((Ba.T @ Ba) * CtC)
    else:
        rhs, cross = intermediates
        inner = sum(np.sum(r * a) for r, a in zip(rhs, A))
        ncmf = sum(np.sum(np.diag(a) @ cross[i] @ np.diag(a)) for i, a in enumerate(A))
    return np.sqrt(max(0, nx2 - 2 * inner + ncmf))


def _l2_term(A, Bs, C, l2):  # decomposition.py:617-627
    out = 0
    if l2[0]:
        out += 0.5 * l2[0] * np.sum(A ** 2)
    if l2[1]:
        out += 0.5 * l2[1] * sum(np.sum(B ** 2) for B in Bs)
    if l2[2]:
        out += 0.5 * l2[2] * np.sum(C ** 2)
    return out


def _feasible(gaps, tol):  # decomposition.py:630-640
    worst = -float("inf")
    for g in gaps:
        if len(g):
            worst = max(max(g), worst)
    return worst < tol


def _penalty_sum(regs, A, Bs, C):
    return (sum(r.value(A) for r in regs[0]) + sum(r.value(Bs) for r in regs[1]) + sum(r.value(C) for r in regs[2]))


# ----------------------------------------------------------------------------------------------------------
# driver (decomposition.py:862-1100)
# ----------------------------------------------------------------------------------------------------------
def ao_admm(matrices, rank, n_iter_max=1000, l2_penalty=None, l1_penalty=None, non_negative=None, unimodal=None,
            l2_norm_bound=None, lower_bound=None, upper_bound=None, parafac2=None, regs=None,
            feasibility_penalty_scale=1, constant_feasibility_penalty=False, aux_init="random_uniform",
            dual_init="random_uniform", random_state=None, tol=1e-8, absolute_tol=1e-10, feasibility_tol=1e-4,
            inner_n_iter_max=5, update_A=True, update_B_is=True, update_C=True, return_errors=True, init=None,
            trajectory=None, inner_tol=None, generalized_l2_penalty=None, tv_penalty=None, regs_spec=None):
    """Returns a dict with factors, auxes, duals and the diagnostics lists of ``return_errors=True``.

    ``trajectory``: optional list; after every outer iteration a dict of deep copies
    (A, B_is, C, auxes, duals) is appended (used to produce / compare per-iteration goldens).
    ``init``: optional (A, B_is, C) to start from instead of the random draw.
    ``regs``: optional explicit [[...],[...],[...]] of OraclePenalty appended after the keyword penalties.
    """
    rs = random_state if isinstance(random_state, np.random.RandomState) else (
        np.random.mtrand._rand if random_state is None else np.random.RandomState(random_state))
    matrices = [np.asarray(M, dtype=np.float64) for M in matrices]
    I, K = len(matrices), matrices[0].shape[1]
    if init is None or (isinstance(init, str) and init == "random"):  # decomposition.py:31-39 — draw order A, C, B_i
        A = rs.uniform(size=(I, rank))
        C = rs.uniform(size=(K, rank))
        Bs = [rs.uniform(size=(M.shape[0], rank)) for M in matrices]
    elif isinstance(init, str) and init in ("svd", "threshold_svd"):  # decomposition.py:42-53
        def truncated(M, n):
            U, S, Vh = np.linalg.svd(M, full_matrices=n > min(M.shape))
            return U[:, :n], S[:n], Vh[:n, :]
        A = np.ones((I, rank))
        Bs = [truncated(M, rank)[0] for M in matrices]
        C = np.transpose(truncated(np.concatenate(matrices, 0), rank)[2])
        if init == "threshold_svd":
            Bs = [np.clip(B, 0, float("inf")) for B in Bs]
            C = np.clip(C, 0, float("inf"))
    else:
        A, Bs, C = np.array(init[0]), [np.array(b) for b in init[1]], np.array(init[2])

    l2 = [v if v is not None else 0 for v in _listify(l2_penalty, "l2_penalty")]
    parsed = build_penalties(non_negative, lower_bound, upper_bound, l2_norm_bound, unimodal, parafac2, l1_penalty,
                             aux_init, dual_init, generalized_l2_penalty, tv_penalty)
    if regs_spec is not None:
        regs = regs_from_spec(regs_spec)
    extra = regs if regs is not None else [[], [], []]
    regs = [parsed[m] + list(extra[m]) for m in range(3)]
    if not update_A:
        regs[0] = []
    if not update_B_is:
        regs[1] = []
    if not update_C:
        regs[2] = []

    aux = [[r.init_aux(matrices, rank, m, rs) for r in regs[m]] for m in range(3)]   # :78-82
    dual = [[r.init_dual(matrices, rank, m, rs) for r in regs[m]] for m in range(3)]  # :85-89
    norm_X = _rss(matrices)
    rec_errors = [reconstruction_error(matrices, A, Bs, C, norm_X) / norm_X]
    losses = [0.5 * rec_errors[0] ** 2 + _l2_term(A, Bs, C, l2) + _penalty_sum(regs, A, Bs, C)]
    gaps = [feasibility_gaps(A, Bs, C, regs, *aux)]

    satisfied, message = False, "MAXIMUM NUMBER OF ITERATIONS REACHED"
    cfp = constant_feasibility_penalty
    if isinstance(cfp, str) and cfp not in {"A", "B"}:
        raise ValueError(f"If `constant_feasibility_penalty` is a string, it must be 'A' or 'B', not {cfp}")
    const_A = (cfp and not isinstance(cfp, str)) or cfp == "A"
    const_B = (cfp and not isinstance(cfp, str)) or cfp == "B"
    feasible = None
    it = -1
    for it in range(n_iter_max):
        inter = None
        if update_B_is:
            Bs, aux[1], dual[1] = solve_mode_B(matrices, regs[1], A, Bs, C, aux[1], dual[1], l2[1], inner_n_iter_max,
                                           feasibility_penalty_scale, const_B, inner_tol)
        if update_C:
            C, aux[2], dual[2] = solve_mode_C(matrices, regs[2], A, Bs, C, aux[2], dual[2], l2[2], inner_n_iter_max,
                                          feasibility_penalty_scale, inner_tol)
        if update_A:
            A, aux[0], dual[0], inter = solve_mode_A(matrices, regs[0], A, Bs, C, aux[0], dual[0], l2[0],
                                                 inner_n_iter_max, feasibility_penalty_scale, const_A, inner_tol)
        if trajectory is not None:
            trajectory.append(snapshot(A, Bs, C, aux, dual))
        if not (tol or absolute_tol or return_errors):  # decomposition.py:990, 1055
            continue
        cur = feasibility_gaps(A, Bs, C, regs, *aux)
        gaps.append(cur)
        if tol or absolute_tol:
            feasible = feasibility_tol and _feasible(cur, feasibility_tol)
            if not feasible and not return_errors:  # loss NOT appended on infeasible iterations (:996-1011)
                continue
        err = reconstruction_error(matrices, A, Bs, C, norm_X, inter) / norm_X
        rec_errors.append(err)
        losses.append(0.5 * err ** 2 + _l2_term(A, Bs, C, l2) + _penalty_sum(regs, A, Bs, C))
        if tol:
            rel_ok = abs(losses[-2] - losses[-1]) < (tol * losses[-2])
            abs_ok = losses[-1] < absolute_tol
            if feasible and rel_ok:
                satisfied, message = True, "FEASIBILITY GAP CRITERION AND RELATIVE LOSS CRITERION SATISFIED"
                break
            elif feasible and abs_ok:
                satisfied, message = True, "FEASIBILITY GAP CRITERION AND ABSOLUTE LOSS CRITERION SATISFIED"
                break
    if feasibility_tol and return_errors:  # decomposition.py:1065-1070
        feasible = _feasible(feasibility_gaps(A, Bs, C, regs, *aux), feasibility_tol)
    elif not feasibility_tol:
        feasible = None
    if not satisfied and not (tol or absolute_tol):
        satisfied = None
    return dict(A=A, B_is=Bs, C=C, aux=aux, dual=dual, regs=regs, rec_errors=rec_errors, regularized_loss=losses,
                feasibility_gaps=gaps, n_iter=it + 1, message=message, satisfied_stopping_condition=satisfied,
                satisfied_feasibility_condition=feasible)


def snapshot(A, Bs, C, aux, dual):
    def cp(v):
        if isinstance(v, np.ndarray):
            return v.copy()
        if isinstance(v, (list, tuple)):
            return type(v)(cp(u) for u in v)
        return v
    return dict(A=A.copy(), B_is=[b.copy() for b in Bs], C=C.copy(), aux=cp(aux), dual=cp(dual))
