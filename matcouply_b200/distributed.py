"""Slice sharding of the AO-ADMM path over the GPUs of one NVLink/NVSwitch box (one process per GPU, NCCL).

The reference is single-process; its cross-slice reductions are Python ``sum`` loops (decomposition.py:310-315 for the
C normal equations, penalties.py:1240-1245 for the PARAFAC2 coordinate matrix, decomposition.py:406-415, 448-449 for
gaps and fit).  Those sites are exactly the all-reduces of the sharded engine (``_engine.AOADMMEngine._allreduce``):

=====================================  ===========================================  ======================
when                                   payload (fp64)                               reference site
=====================================  ===========================================  ======================
once, before the first iteration       ``||X||^2``                                  decomposition.py:906
C-update, once per outer iteration     ``Z`` (K x R) + ``lhs_C`` (R x R), one buffer  decomposition.py:310-315
each B inner iteration, PARAFAC2 only  ``sum rho_i P_i^T V_i`` (R x R) + ``sum rho_i``  penalties.py:1240-1245
constant feasibility penalty           MAX of rho (1 scalar) for mode A / mode B    decomposition.py:164, 249
end of outer iteration                 packed gap / fit / norm scalars              decomposition.py:406-415, 448-449
=====================================  ===========================================  ======================

Everything indexed by slice (``X_i``, ``B_i``, rows ``a_i``, their aux/dual variables, per-slice operators) stays
rank-local; ``C``, ``Delta`` and all scalars are replicated (every rank computes identical values from identical
reduced inputs, so the host stopping decision is identical on all ranks without a broadcast).

This module holds the host-side logic: the row-balanced slice partition and the extraction of a rank's share of a
globally drawn initial state (the reference draws A, C, B_0.., then every aux and dual, from ONE ``RandomState`` —
decomposition.py:31-39, 78-89 — so bit-parity with an unsharded run needs the global draw on every rank).
"""
from typing import List, NamedTuple, Sequence, Tuple

import numpy as np


class ShardSpec(NamedTuple):
    """This rank's contiguous range ``[lo, hi)`` of the global slice list and the global slice heights J_i."""

    row_counts: Tuple[int, ...]  # J_i of EVERY slice of the global problem
    lo: int
    hi: int
    n_cols: int = 0  # K of the data matrices; only needed by a rank whose range is EMPTY (it has no matrix to ask)
    dtype: str = "float64"  # "float32" selects the fp32 kernels on such a rank (the others see it from their data)

    @property
    def n_global(self):
        return len(self.row_counts)


def partition_slices(row_counts: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous slice ranges balanced by ROW count (the X-stream cost is proportional to sum J_i, not to the number
    of slices; BASELINE config 2 is ragged 8x).  Every rank gets a (possibly empty) range; ranges tile [0, I)."""
    sizes = np.asarray(row_counts, dtype=np.int64)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world_size):
        cut = int(np.searchsorted(csum, total * r / world_size))
        cuts.append(min(max(cut, cuts[-1]), len(sizes)))
    cuts.append(len(sizes))
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def make_shard(row_counts: Sequence[int], rank: int, world_size: int, n_cols: int = 0,
               dtype: str = "float64") -> ShardSpec:
    """``n_cols`` / ``dtype``: the column count K and the element type of the data; pass them when ``world_size`` can
    exceed the number of slices, so that a rank with an empty range still knows the shape of the problem."""
    lo, hi = partition_slices(row_counts, world_size)[rank]
    return ShardSpec(tuple(int(j) for j in row_counts), lo, hi, int(n_cols), str(dtype))


def shard_state(A, B_is, auxes, duals, regs, shard: ShardSpec):
    """Cut this rank's share out of a globally initialised state.

    ``A`` (I x R), ``B_is`` (list of I matrices), ``auxes`` / ``duals`` (3 lists, one entry per penalty, in the
    reference's ADMMVars layout).  Mode 0 variables are row-sliced, mode 1 lists are list-sliced (a PARAFAC2 aux is the
    tuple ``(basis list, Delta)``: the bases are sliced, Delta is replicated), mode 2 variables are replicated."""
    lo, hi = shard.lo, shard.hi

    def rows(v):
        return v.cut(lo, hi) if hasattr(v, "cut") else list(v[lo:hi])  # DeviceRows stay on the device

    def cut(mode, v):
        if mode == 0:
            return np.asarray(v)[lo:hi]
        if mode == 1:
            if isinstance(v, tuple):  # PARAFAC2: (bases, coordinate matrix)
                return (rows(v[0]), v[1])
            return rows(v)
        return v

    A_loc = np.asarray(A)[lo:hi]
    B_loc = rows(B_is)
    aux_loc = [[cut(m, v) for v in auxes[m]] for m in range(3)]
    dual_loc = [[cut(m, v) for v in duals[m]] for m in range(3)]
    return A_loc, B_loc, aux_loc, dual_loc


def allgather_rows(local, shard: ShardSpec, group):
    """All-gather of a row-sharded host array list (used to hand every rank the complete A / B_is on return).
    ``local`` is a list of NumPy arrays (one per local slice) or a 2-D array whose rows are the local slices."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, local, group=group)
    if isinstance(local, np.ndarray):
        return np.concatenate(out, 0)
    merged = []
    for part in out:
        merged.extend(part)
    return merged
