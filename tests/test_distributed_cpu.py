"""Host-side logic of the slice-sharded engine, checked on CPU with a world_size-2 ``gloo`` group.

The CUDA engine itself cannot run here (no GPU); what CAN be checked without one is everything the sharding adds on
top of it: the row-balanced partition, the cut of a globally drawn initial state (RNG parity with the unsharded run),
and that the reference's cross-slice sums decompose exactly into (rank-local partial, one all-reduce of the packed
payload) at each of the reduction sites listed in ``matcouply_b200/distributed.py``.  The N>1 CUDA path proper is
covered by ``tests/test_gpu_multi.py`` (needs 2 GPUs).
"""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from matcouply_b200 import decomposition as D  # noqa: E402
from matcouply_b200.distributed import make_shard, partition_slices, shard_state  # noqa: E402


def test_partition_tiles_and_balances_rows():
    rs = np.random.RandomState(0)
    for world in (1, 2, 3, 8):
        sizes = rs.randint(256, 2049, size=257)
        parts = partition_slices(sizes, world)
        assert parts[0][0] == 0 and parts[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
        rows = [int(sizes[lo:hi].sum()) for lo, hi in parts]
        assert max(rows) - min(rows) <= 2 * 2048  # within two slices of perfect balance
    assert partition_slices([5], 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or sum(
        hi - lo for lo, hi in partition_slices([5], 4)) == 1
    with pytest.raises(ValueError):
        partition_slices([1, 2], 0)


def _global_state(row_counts, K, R, kw, seed=0):
    shapes = [D._ShapeOnly((j, K)) for j in row_counts]
    rs = np.random.RandomState(seed)
    cmf = D.initialize_cmf(shapes, R, "random", random_state=rs)
    regs = D._parse_all_penalties(
        non_negative=kw.get("non_negative"), lower_bound=None, upper_bound=None, l2_norm_bound=kw.get("l2_norm_bound"),
        unimodal=kw.get("unimodal"), parafac2=kw.get("parafac2"), l1_penalty=kw.get("l1_penalty"),
        tv_penalty=kw.get("tv_penalty"), generalized_l2_penalty=kw.get("generalized_l2_penalty"), svd="truncated_svd",
        regs=kw.get("regs"), dual_init="random_uniform",
        aux_init="random_uniform", verbose=False)
    auxes = [[r.init_aux(shapes, R, m, random_state=rs) for r in regs[m]] for m in range(3)]
    duals = [[r.init_dual(shapes, R, m, random_state=rs) for r in regs[m]] for m in range(3)]
    return cmf, regs, auxes, duals


def _shard_cases():
    from matcouply_b200 import penalties as P

    lap = 2 * np.eye(11) - np.eye(11, k=1) - np.eye(11, k=-1)
    return [dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}),
            # the penalties added later: TV (mode 1), generalized L2 (mode 2), unit simplex (mode 0), PARAFAC2 options
            dict(non_negative={2: True}, tv_penalty={1: 0.1}, generalized_l2_penalty={2: lap},
                 regs=[[P.UnitSimplex()], [P.Parafac2(n_iter=2)], []])]


@pytest.mark.parametrize("case", [0, 1])
def test_shard_state_reassembles_global_draw(case):
    row_counts, K, R = [7, 3, 9, 4, 6, 8, 5], 11, 3
    kw = _shard_cases()[case]
    cmf, regs, auxes, duals = _global_state(row_counts, K, R, kw)
    _, (A, B_is, C) = cmf
    world = 3
    parts = [shard_state(A, B_is, auxes, duals, regs, make_shard(row_counts, r, world)) for r in range(world)]
    np.testing.assert_array_equal(np.concatenate([p[0] for p in parts], 0), A)
    flat = [b for p in parts for b in p[1]]
    assert len(flat) == len(B_is) and all(np.array_equal(x, y) for x, y in zip(flat, B_is))
    for m in range(3):
        for k in range(len(regs[m])):
            for which, glob in ((2, auxes), (3, duals)):
                pieces = [p[which][m][k] for p in parts]
                g = glob[m][k]
                if m == 2:
                    assert all(np.array_equal(x, g) for x in pieces)  # replicated
                elif m == 0:
                    np.testing.assert_array_equal(np.concatenate(pieces, 0), g)
                elif isinstance(g, tuple):  # PARAFAC2 aux: bases sharded, Delta replicated
                    assert all(np.array_equal(x[1], g[1]) for x in pieces)
                    bases = [b for x in pieces for b in x[0]]
                    assert all(np.array_equal(x, y) for x, y in zip(bases, g[0]))
                else:
                    fl = [b for x in pieces for b in x]
                    assert all(np.array_equal(x, y) for x, y in zip(fl, g))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, row_counts, K, R, out_dir):
    """One logical rank: NumPy partials of every reduction site on its shard + the engine's packed all-reduces."""
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(1)
        X = [rs.standard_normal(size=(j, K)) for j in row_counts]          # same global data on every rank
        cmf, regs, auxes, duals = _global_state(row_counts, K, R, dict(non_negative=True, parafac2=True))
        _, (A, B_is, C) = cmf
        sh = make_shard(row_counts, rank, world)
        A_l, B_l, aux_l, dual_l = shard_state(A, B_is, auxes, duals, regs, sh)
        X_l = X[sh.lo:sh.hi]

        # (1) ||X||^2  (decomposition.py:906)
        nx = torch.tensor([sum(float(np.sum(x ** 2)) for x in X_l)], dtype=torch.float64)
        dist.all_reduce(nx)
        # (2) C normal equations: Z (K x R) and lhs_C (R x R) in ONE buffer, like AOADMMEngine.ZL  (:310-315)
        ZL = torch.zeros(K * R + R * R, dtype=torch.float64)
        Z, lhs = ZL[: K * R].view(K, R).numpy(), ZL[K * R:].view(R, R).numpy()
        for x, b, a in zip(X_l, B_l, A_l):
            ba = b * a
            Z += x.T @ ba
            lhs += ba.T @ ba
        dist.all_reduce(ZL)
        # (3) PARAFAC2 coordinate matrix: sum rho_i P_i^T V_i (R x R) and sum rho_i   (penalties.py:1240-1245)
        sums = torch.zeros(R * R + 1, dtype=torch.float64)
        num, Delta = sums[: R * R].view(R, R).numpy(), aux_l[1][0][1]
        for i, (b, d) in enumerate(zip(B_l, dual_l[1][0])):
            V = b + d
            U, _, Vh = np.linalg.svd(V @ Delta.T, full_matrices=False)
            rho_i = 1.0 + 0.1 * (sh.lo + i)
            num += rho_i * (U @ Vh).T @ V
            sums[R * R] += rho_i
        dist.all_reduce(sums)
        # (4) constant feasibility penalty: MAX over slices  (decomposition.py:164, 249)
        rmax = torch.tensor([max([float(np.sum(a ** 2)) for a in A_l], default=-np.inf)], dtype=torch.float64)
        dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
        # (5) gap / norm scalars of the sharded modes  (:406-415)
        scal = torch.tensor([float(np.sum((A_l - aux_l[0][0]) ** 2)), float(np.sum(A_l ** 2)),
                             sum(float(np.sum(b ** 2)) for b in B_l)], dtype=torch.float64)
        dist.all_reduce(scal)
        if rank == 0:
            np.savez(os.path.join(out_dir, "reduced.npz"), nx=nx.numpy(), ZL=ZL.numpy(), sums=sums.numpy(),
                     rmax=rmax.numpy(), scal=scal.numpy())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_reduction_sites_equal_unsharded(tmp_path):
    import torch.multiprocessing as mp

    row_counts, K, R, world = [6, 9, 4, 7, 5, 8], 10, 3, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, row_counts, K, R, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "reduced.npz"))

    rs = np.random.RandomState(1)
    X = [rs.standard_normal(size=(j, K)) for j in row_counts]
    cmf, regs, auxes, duals = _global_state(row_counts, K, R, dict(non_negative=True, parafac2=True))
    _, (A, B_is, C) = cmf
    np.testing.assert_allclose(got["nx"][0], sum(np.sum(x ** 2) for x in X), rtol=1e-13)
    Z = sum(x.T @ (b * a) for x, b, a in zip(X, B_is, A))
    lhs = sum((b * a).T @ (b * a) for b, a in zip(B_is, A))
    np.testing.assert_allclose(got["ZL"][: K * R].reshape(K, R), Z, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got["ZL"][K * R:].reshape(R, R), lhs, rtol=1e-12)
    Delta = auxes[1][0][1]
    num, den = np.zeros((R, R)), 0.0
    for i, (b, d) in enumerate(zip(B_is, duals[1][0])):
        V = b + d
        U, _, Vh = np.linalg.svd(V @ Delta.T, full_matrices=False)
        num += (1.0 + 0.1 * i) * (U @ Vh).T @ V
        den += 1.0 + 0.1 * i
    np.testing.assert_allclose(got["sums"][: R * R].reshape(R, R) / got["sums"][R * R], num / den, rtol=1e-12)
    np.testing.assert_allclose(got["rmax"][0], max(np.sum(a ** 2) for a in A), rtol=1e-15)
    np.testing.assert_allclose(got["scal"], [np.sum((A - auxes[0][0]) ** 2), np.sum(A ** 2),
                                             sum(np.sum(b ** 2) for b in B_is)], rtol=1e-13)
