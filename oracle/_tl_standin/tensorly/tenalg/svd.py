import numpy as np


def truncated_svd(matrix, n_eigenvecs=None, **kwargs):
    d1, d2 = matrix.shape
    min_dim, max_dim = min(d1, d2), max(d1, d2)
    if n_eigenvecs is None:
        n_eigenvecs = max_dim
    U, S, V = np.linalg.svd(matrix, full_matrices=n_eigenvecs > min_dim)
    return U[:, :n_eigenvecs], S[:n_eigenvecs], V[:n_eigenvecs, :]
