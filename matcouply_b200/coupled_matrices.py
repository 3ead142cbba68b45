"""Result container of the decomposition: same behaviour as ``matcouply.coupled_matrices.CoupledMatrixFactorization``
(reference src/matcouply/coupled_matrices.py:8-240) for the parts the AO-ADMM path returns and consumes: tuple-like
``(weights, (A, B_is, C))``, ``.shape``, ``.rank``, validation (``_validate_cmf`` :243-362) and dense reconstruction
(``cmf_to_matrix`` :365-428, ``cmf_to_matrices`` :441-504).  Host-side convenience on NumPy arrays (or torch tensors);
the TensorLy conversions (``from_CPTensor``, ``to_tensor`` ...) are out of scope.
"""
import numpy as np

__all__ = ["CoupledMatrixFactorization", "cmf_to_matrix", "cmf_to_matrices", "cmf_to_slice", "cmf_to_slices"]


def _is_array(x):
    if isinstance(x, np.ndarray):
        return True
    try:
        import torch

        return isinstance(x, torch.Tensor)
    except ImportError:  # pragma: no cover
        return False


def _validate_cmf(cmf):
    weights, (A, B_is, C) = cmf
    if not (_is_array(weights) or weights is None):
        raise TypeError("Weights should be a first order tensor of length rank, not {}".format(type(weights)))
    elif weights is not None and len(weights.shape) != 1:
        raise ValueError(
            "Weights should be a first order tensor. However weights has shape {}".format(tuple(weights.shape))
        )
    for label, position, sizes, M in (("A", "first", "(I, rank)", A), ("C", "last", "(K, rank)", C)):
        if not _is_array(M):
            raise TypeError(
                "The {} factor matrix, {}, should be a second order tensor of size {}), not {}".format(
                    position, label, sizes, type(M)
                )
            )
        elif len(M.shape) != 2:
            raise ValueError(
                "The {} factor matrix, {}, should be a second order tensor. However {} has shape {}".format(
                    position, label, label, tuple(M.shape)
                )
            )
    rank = int(A.shape[1])
    if C.shape[1] != rank:
        raise ValueError(
            "All the factors of a coupled matrix factorization should have the same number of columns."
            "However, A.shape[1]={} but C.shape[1]={}.".format(rank, C.shape[1])
        )
    shape = []
    for i, B_i in enumerate(B_is):
        if not _is_array(B_i):
            raise TypeError(
                "The B_is[{}] factor matrix should be second order tensor of size (J_i, rank)), not {}".format(
                    i, type(B_i)
                )
            )
        elif len(B_i.shape) != 2:
            raise ValueError(
                "The B_is[{}] factor matrix should be second order tensor. However B_is[{}] has shape {}".format(
                    i, i, tuple(B_i.shape)
                )
            )
        if B_i.shape[1] != rank:
            raise ValueError(
                "All the factors of a coupled matrix factorization should have the same number of columns."
                "However, A.shape[1]={} but B_is[{}].shape[1]={}.".format(rank, i, B_i.shape[1])
            )
        shape.append((B_i.shape[0], C.shape[0]))
    if weights is not None and weights.shape[0] != rank:
        raise ValueError(
            "Given factors for a rank-{} coupled matrix factorization but len(weights)={}.".format(
                rank, weights.shape[0]
            )
        )
    if A.shape[0] != len(B_is):
        raise ValueError(
            "The number of rows in A should be the same as the number of B_i matrices"
            "However, tl.shape(A)[0]={}, but len(B_is)={}".format(A.shape[0], len(B_is))
        )
    return tuple(shape), rank


class CoupledMatrixFactorization:
    """``(weights, (A, B_is, C))`` with ``X_i ~ B_i diag(a_i) C^T``; indexable and iterable like a 2-tuple."""

    def __init__(self, cmf_matrices):
        shape, rank = _validate_cmf(cmf_matrices)
        self.weights, self.factors = cmf_matrices
        self.shape = shape
        self.rank = rank

    @classmethod
    def from_CPTensor(cls, cp_tensor, shapes=None):
        """A third-order CP tensor ``(weights, (A, B, C))`` (any 2-tuple: TensorLy ``CPTensor`` objects unpack the same
        way) as a coupled matrix factorization with ``B_i = B[:J_i]`` (coupled_matrices.py:101-151)."""
        weights, factors = cp_tensor
        if len(factors) != 3:
            raise ValueError("Must be a third order CP tensor to convert into a coupled matrix factorization")
        A, B, C = (np.asarray(f) for f in factors)
        if shapes is not None:
            B_is = []
            if len(shapes) != A.shape[0]:
                raise ValueError(
                    f"The first mode has length {A.shape[0]}, which is different "
                    f"than the length indicated by the shapes argument ({len(shapes)})"
                )
            for J_i, K in shapes:
                if K != C.shape[0]:
                    raise ValueError(
                        f"The third mode has length {C.shape[0]}, which is different "
                        f"than the length indicated by the shapes argument ({K})"
                    )
                if J_i > B.shape[0]:
                    raise ValueError(
                        f"The second mode of the CP tensor mode has length {B.shape[0]}, which "
                        f"is smaller than the length indicated by the shape ({J_i}) of matrix"
                    )
                B_is.append(np.copy(B)[:J_i, :])
        else:
            B_is = [np.copy(B) for _ in range(A.shape[0])]
        weights = np.ones(A.shape[1]) if weights is None else np.copy(weights)
        return cls((weights, [np.copy(A), B_is, np.copy(C)]))

    @classmethod
    def from_Parafac2Tensor(cls, parafac2_tensor):
        """A PARAFAC2 tensor ``(weights, (A, B, C), projection_matrices)`` as a coupled matrix factorization with
        ``B_i = P_i B`` (coupled_matrices.py:153-172)."""
        weights, factors, projection_matrices = parafac2_tensor
        A, B, C = (np.asarray(f) for f in factors)
        B_is = [np.dot(np.asarray(P_i), B) for P_i in projection_matrices]
        weights = np.ones(A.shape[1]) if weights is None else np.copy(weights)
        return cls((weights, [np.copy(A), B_is, np.copy(C)]))

    def __getitem__(self, item):
        if item == 0:
            return self.weights
        elif item == 1:
            return self.factors
        raise IndexError(
            "You tried to access index {} of a coupled matrix factorization.\n"
            "You can only access index 0 and 1 of a coupled matrix factorization"
            "(corresponding respectively to the weights and factors)".format(item)
        )

    def __iter__(self):
        yield self.weights
        yield self.factors

    def __len__(self):
        return 2

    def __repr__(self):  # pragma: nocover
        return "(weights, factors) : rank-{} CoupledMatrixFactorization of shape {}".format(self.rank, self.shape)

    def to_tensor(self):
        return cmf_to_tensor(self)

    def to_vec(self, pad=True):
        return cmf_to_vec(self, pad=pad)

    def to_unfolded(self, mode, pad=True):
        return cmf_to_unfolded(self, mode, pad=pad)

    def to_matrices(self):
        return cmf_to_matrices(self)

    def to_matrix(self, matrix_idx):
        return cmf_to_matrix(self, matrix_idx)


def cmf_to_matrix(cmf, matrix_idx, validate=True):
    if validate:
        cmf = CoupledMatrixFactorization(cmf)
    weights, (A, B_is, C) = cmf
    a = A[matrix_idx]
    if weights is not None:
        a = a * weights
    return (B_is[matrix_idx] * a) @ C.T


def cmf_to_matrices(cmf, validate=True):
    if validate:
        cmf = CoupledMatrixFactorization(cmf)
    weights, (A, B_is, C) = cmf
    if weights is not None:
        A = A * weights
    return [cmf_to_matrix((None, (A, B_is, C)), i, validate=False) for i in range(A.shape[0])]


def cmf_to_slice(cmf, slice_idx, validate=True):
    """Alias of ``cmf_to_matrix`` with the reference's keyword name (coupled_matrices.py:431)."""
    return cmf_to_matrix(cmf, slice_idx, validate=validate)


def cmf_to_slices(cmf, validate=True):
    """Alias of ``cmf_to_matrices`` (coupled_matrices.py:507)."""
    return cmf_to_matrices(cmf, validate=validate)


def cmf_to_tensor(cmf, validate=True):
    """Dense ``I x max(J_i) x K`` tensor, shorter matrices zero-padded at the bottom (coupled_matrices.py:517-596)."""
    _, (A, B_is, C) = cmf
    matrices = cmf_to_matrices(cmf, validate=validate)
    lengths = [B_i.shape[0] for B_i in B_is]
    tensor = np.zeros((A.shape[0], max(lengths), C.shape[0]), dtype=np.asarray(matrices[0]).dtype)
    for i, (matrix, length) in enumerate(zip(matrices, lengths)):
        tensor[i, :length] = matrix
    return tensor


def cmf_to_unfolded(cmf, mode, pad=True, validate=True):
    """Mode-``mode`` unfolding of the (padded) tensor; ``pad=False`` only for mode 2 (coupled_matrices.py:599-714)."""
    if pad:
        tensor = cmf_to_tensor(cmf, validate=validate)
        return np.reshape(np.moveaxis(tensor, mode, 0), (tensor.shape[mode], -1))
    if mode == 2:
        return np.transpose(np.concatenate(cmf_to_matrices(cmf, validate=validate), axis=0))
    raise ValueError(f"Cannot unfold along mode {mode} without padding. ")


def cmf_to_vec(cmf, pad=True, validate=True):
    """Vectorised (padded) tensor, or the concatenated raveled matrices (coupled_matrices.py:717-799)."""
    if pad:
        return np.reshape(cmf_to_tensor(cmf, validate=validate), (-1,))
    return np.concatenate([np.reshape(m, (-1,)) for m in cmf_to_matrices(cmf, validate=validate)])
