"""Minimal NumPy stand-in for the third-party `tensorly` package (NOT installed in this image).

TEST INFRASTRUCTURE ONLY.  It lets the unmodified reference (`/root/reference/src/matcouply`) be
imported in the build container so that `oracle/gen_golden.py` can produce golden vectors and
`tests/test_oracle_vs_reference.py` can pin the oracle.  It is a dependency substitute (every name maps
1:1 onto NumPy, exactly like TensorLy's own NumPy backend); it contains no AO-ADMM logic.
"""
import numpy as np

from . import tenalg, _factorized_tensor, testing, decomposition  # noqa: F401


def get_backend():
    return "numpy"


def tensor(x, **kw):
    return np.array(x, **kw)


def is_tensor(x):
    return isinstance(x, np.ndarray)


def shape(x):
    return x.shape


def to_numpy(x):
    return np.asarray(x)


def context(x):
    return {"dtype": x.dtype}


def check_random_state(seed):
    if seed is None:
        return np.random.mtrand._rand
    if isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError("Seed should be None, int or np.random.RandomState")


dot = np.dot
matmul = np.matmul
transpose = np.transpose
sum = np.sum
abs = np.abs
sign = np.sign
sqrt = np.sqrt
clip = np.clip
zeros = np.zeros
ones = np.ones
eye = np.eye
diag = np.diag
trace = np.trace
copy = np.copy
stack = np.stack
concatenate = np.concatenate
all = np.all
min = np.min
max = np.max
reshape = np.reshape
zeros_like = np.zeros_like
solve = np.linalg.solve


def norm(x, order=2, axis=None):
    return np.sqrt(np.sum(np.abs(x) ** 2, axis=axis))


class _Index:
    def __getitem__(self, k):
        return k


index = _Index()


def index_update(t, idx, val):
    t[idx] = val
    return t


SVD_FUNS = ["truncated_svd", "symeig_svd", "randomized_svd"]


def unfold(t, mode):  # tensorly.base.unfold: mode-`mode` unfolding, remaining modes in C order
    return np.reshape(np.moveaxis(t, mode, 0), (t.shape[mode], -1))


def tensor_to_vec(t):  # tensorly.base.tensor_to_vec
    return np.reshape(t, (-1,))


class _CPTensor:
    """tensorly.cp_tensor.CPTensor as the reference uses it (coupled_matrices.py:121-122): unpacks as
    (weights, factors), weights defaulting to ones(rank)."""

    def __init__(self, cp_tensor):
        weights, factors = cp_tensor
        rank = np.shape(factors[0])[1]
        self.weights = np.ones(rank) if weights is None else weights
        self.factors = factors

    def __iter__(self):
        yield self.weights
        yield self.factors


class _Parafac2Tensor:
    """tensorly.parafac2_tensor.Parafac2Tensor (coupled_matrices.py:168-169): (weights, factors, projections)."""

    def __init__(self, parafac2_tensor):
        weights, factors, projections = parafac2_tensor
        rank = np.shape(factors[0])[1]
        self.weights = np.ones(rank) if weights is None else weights
        self.factors, self.projections = factors, projections

    def __iter__(self):
        yield self.weights
        yield self.factors
        yield self.projections


class cp_tensor:  # namespace stand-ins for the two sub-modules
    CPTensor = _CPTensor


class parafac2_tensor:
    Parafac2Tensor = _Parafac2Tensor
