"""matcouply_b200 — B200-native AO-ADMM engine behind MatCoupLy's ``cmf_aoadmm`` API.

Public modules mirror the reference package (``matcouply/__init__.py:10``):
``decomposition`` (cmf_aoadmm, parafac2_aoadmm, ...), ``penalties`` (ADMMPenalty subclasses) and
``coupled_matrices`` (CoupledMatrixFactorization).  All numerics run as hand-written sm_100a CUDA kernels from
``matcouply_b200/csrc`` behind the C ABI in ``include/matcouply_b200.h``; importing the package does not need a GPU,
calling it does.
"""
from . import coupled_matrices, data, decomposition, penalties, random  # noqa: F401
from .coupled_matrices import CoupledMatrixFactorization  # noqa: F401
from .decomposition import ADMMVars, DiagnosticMetrics, cmf_aoadmm, parafac2_aoadmm  # noqa: F401

__version__ = "0.1.0"
__all__ = ["coupled_matrices", "data", "decomposition", "penalties", "random", "CoupledMatrixFactorization", "cmf_aoadmm",
           "parafac2_aoadmm", "ADMMVars", "DiagnosticMetrics", "PackedMatrices"]


def __getattr__(name):  # torch is imported lazily so that `import matcouply_b200` stays light
    if name == "PackedMatrices":
        from ._engine import PackedMatrices

        return PackedMatrices
    raise AttributeError(name)
