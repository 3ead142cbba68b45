"""``unimodal_regression`` with the reference's signature (``matcouply/_unimodal_regression.py:107-141``), as custom
penalties import it (``examples/plot_custom_penalty.py:214``).  The projection runs in the CUDA kernel behind
``Unimodality`` (``csrc/unimodal.cu``: prefix-isotonic regression from both ends + first strict minimum of the summed
errors, bit-exact peak index against the reference's operation order); there is no CPU implementation here."""
import numpy as np

__all__ = ["unimodal_regression"]


def unimodal_regression(y, non_negativity=False):
    """Project ``y`` (a vector, or an N-d array whose first-mode fibers are projected) onto the unimodal vectors,
    optionally non-negative.  Returns a NumPy array of the same shape."""
    from .penalties import Unimodality

    y = np.asarray(y)
    if y.size == 0:
        return np.array(y, dtype=np.float64)
    columns = np.ascontiguousarray(y, dtype=np.float32 if y.dtype == np.float32 else np.float64)
    columns = columns.reshape(y.shape[0], -1)
    out = Unimodality(non_negativity=non_negativity).factor_matrix_update(columns, 1.0, None)
    return np.asarray(out).reshape(y.shape)
