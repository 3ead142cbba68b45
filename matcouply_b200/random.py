"""Random coupled matrix factorizations (reference: src/matcouply/random.py:9-66).  Host-side test-data helper."""
import numpy as np

from .coupled_matrices import CoupledMatrixFactorization
from .penalties import _check_random_state


def random_coupled_matrices(shapes, rank, full=False, random_state=None, normalise_factors=True, normalise_B=False,
                            **context):
    """Uniform [0, 1) factors drawn in the reference's order (A, B_0..B_{I-1}, C); with ``normalise_factors`` the column
    norms move into ``weights`` (random.py:40-66).  ``full=True`` returns the dense matrices instead."""
    rns = _check_random_state(random_state)
    if not all(shape[1] == shapes[0][1] for shape in shapes):
        raise ValueError("All matrices must have equal number of columns.")
    dtype = context.get("dtype", np.float64)
    A = np.asarray(rns.random_sample((len(shapes), rank)), dtype=dtype)
    B_is = [np.asarray(rns.random_sample((j_i, rank)), dtype=dtype) for j_i, _k in shapes]
    C = np.asarray(rns.random_sample((shapes[0][1], rank)), dtype=dtype)
    weights = np.ones(rank, dtype=dtype)
    if normalise_factors or normalise_B:
        B_i_norms = [np.sqrt(np.sum(np.abs(B_i) ** 2, axis=0)) for B_i in B_is]
        B_is = [B_i / n for B_i, n in zip(B_is, B_i_norms)]
        A = A * np.stack(B_i_norms)
    if normalise_factors:
        A_norm = np.sqrt(np.sum(np.abs(A) ** 2, axis=0))
        A = A / A_norm
        C_norm = np.sqrt(np.sum(np.abs(C) ** 2, axis=0))
        C = C / C_norm
        weights = A_norm * C_norm
    cmf = CoupledMatrixFactorization((weights, (A, B_is, C)))
    return cmf.to_matrices() if full else cmf
