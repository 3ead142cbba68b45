"""Stand-in for the reference's optional, un-vendored GPL dependency ``condat_tv`` (reference penalties.py:6-11),
which cannot be installed offline.  TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py so that the reference's own
TotalVariationPenalty / cmf_aoadmm code can run): ``tv_denoise_matrix`` denoises every ROW of the matrix with the
oracle's restatement of Condat's direct algorithm (oracle/aoadmm_oracle.py::tv_denoise_1d).  The TV minimiser is
unique, so any exact solver is interchangeable here up to round-off."""
import numpy as np

from oracle.aoadmm_oracle import tv_denoise_1d


def tv_denoise(signal, regularisation_strength):
    return tv_denoise_1d(np.asarray(signal, dtype=np.float64), float(regularisation_strength))


def tv_denoise_matrix(matrix, regularisation_strength):
    matrix = np.asarray(matrix, dtype=np.float64)
    return np.stack([tv_denoise_1d(row, float(regularisation_strength)) for row in matrix], axis=0)
