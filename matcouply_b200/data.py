"""Example data (reference: src/matcouply/data.py).  Only the simulated data set of the README / BASELINE config 0 is
provided; the bike-sharing and semiconductor-etch loaders are file / network I/O and out of scope."""
import numpy as np

from .coupled_matrices import CoupledMatrixFactorization
from .penalties import _check_random_state


def _norm_pdf(t, loc):
    return np.exp(-((t - loc) ** 2) / 2.0) / np.sqrt(2.0 * np.pi)


def get_simple_simulated_data(noise_level=0.2, random_state=1):
    """15 noisy 50 x 20 matrices from a rank-3 model: uniform A (+0.1), cyclically shifted Gaussian bumps as B_i,
    truncated-normal C (data.py:28-95; same draws from ``random_state`` in the same order)."""
    rank = 3
    I, J, K = 15, 50, 20
    rng = _check_random_state(random_state)
    A = rng.uniform(size=(I, rank)) + 0.1
    t = np.linspace(-10, 10, J)
    blueprint = np.stack([_norm_pdf(t, -5), _norm_pdf(t, 0), _norm_pdf(t, 2)], axis=-1)
    B_is = [np.roll(blueprint, i, axis=0) for i in range(I)]
    C = rng.standard_normal(size=(K, rank))
    C[C < 0] = 0
    cmf = CoupledMatrixFactorization((None, (A, B_is, C)))
    matrices = cmf.to_matrices()
    noise = [rng.standard_normal(size=M.shape) for M in matrices]
    scale_factor = np.sqrt(np.sum(np.stack(matrices) ** 2)) / np.sqrt(np.sum(np.stack(noise) ** 2))
    return [M + noise_level * scale_factor * N for M, N in zip(matrices, noise)], cmf
