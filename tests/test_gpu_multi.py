"""N>1 CUDA path: the slice-sharded engine (one process per rank) must reproduce the reference trajectory exactly like
the single-GPU run does, at 2, 4 and 8 ranks.

With enough GPUs every rank owns one and the collectives are NCCL (`gpurun --gpus 8 -- python -m pytest
tests/test_gpu_multi.py -m gpu`).  On a box with fewer GPUs than ranks the SAME sharded CUDA path runs with all ranks
on GPU 0 and `gloo` carrying the (CUDA-tensor) all-reduces — NCCL refuses two ranks on one device — so the sharding
logic, the per-rank kernels and the reduction sites are covered by the driver's one-GPU test run too."""
import json
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _load_case(name):
    g = np.load(os.path.join(HERE, "golden", f"traj_{name}.npz"), allow_pickle=False)
    off = g["row_offsets"]
    X = [g["X"][a:b] for a, b in zip(off[:-1], off[1:])]
    kw = json.loads(str(g["kwargs"]))
    for key in ("l1_penalty", "non_negative", "unimodal", "l2_norm_bound", "lower_bound", "upper_bound", "tv_penalty"):
        if isinstance(kw.get(key), dict):
            kw[key] = {int(k): v for k, v in kw[key].items()}
    if isinstance(kw.get("generalized_l2_penalty"), dict):
        kw["generalized_l2_penalty"] = {int(k): np.asarray(v, dtype=np.float64)
                                        for k, v in kw["generalized_l2_penalty"].items()}
    if "regs_spec" in kw:
        from matcouply_b200 import penalties as P
        from oracle.aoadmm_oracle import regs_from_spec

        kw["regs"] = regs_from_spec(kw.pop("regs_spec"), P)
    return g, X, int(g["rank"]), kw


def _init_group(rank, world, port):
    """NCCL with one GPU per rank when the box has them, else every rank on GPU 0 with gloo (see module docstring)."""
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if torch.cuda.device_count() >= world:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def _backend(world):
    return "nccl" if torch.cuda.device_count() >= world else "gloo, all ranks on GPU 0"


def _worker(rank, world, port, name, k_iter, out_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.distributed import make_shard

    dist = _init_group(rank, world, port)
    try:
        g, X, R, kw = _load_case(name)
        kw = dict(kw)
        kw.pop("n_iter_max", None)
        sh = make_shard([x.shape[0] for x in X], rank, world, n_cols=X[0].shape[1])
        cmf, diag = cmf_aoadmm(X[sh.lo:sh.hi], R, n_iter_max=k_iter, tol=None, absolute_tol=None, return_errors=True,
                               process_group=dist.group.WORLD, shard=sh, **kw)
        _, (A, B_is, C) = cmf
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), A=A, B=np.concatenate(B_is, 0), C=C,
                 loss=np.asarray(diag.regularized_loss), rec=np.asarray(diag.rec_errors))
    finally:
        dist.destroy_process_group()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", ["c2_nn_pf2_l1_ragged", "c1_nn_cmf", "c3_unimodal_l2ball_pf2", "c0_readme",
                                  "gl2_smooth_B_pf2", "tv_B_ragged_constB", "simplex_B_ragged", "pf2_n_iter3_nn",
                                  "pf2_frozen_basis", "tv_C_l1"])
def test_two_rank_run_matches_reference_trajectory(name, tmp_path):
    import torch.multiprocessing as mp

    g, X, R, kw = _load_case(name)
    n_traj = g["A_traj"].shape[0]
    k = min(20, n_traj)
    mp.spawn(_worker, args=(2, _free_port(), name, k, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    for key in ("A", "B", "C", "loss", "rec"):  # gathered factors and diagnostics are identical on both ranks
        np.testing.assert_array_equal(r0[key], r1[key])
    errs = (_rel(r0["A"], g["A_traj"][k - 1]), _rel(r0["B"], g["B_traj"][k - 1]), _rel(r0["C"], g["C_traj"][k - 1]))
    assert max(errs) < 1e-8, errs
    np.testing.assert_allclose(r0["loss"], g["regularized_loss"][: k + 1], rtol=1e-8)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("name", ["c2_nn_pf2_l1_ragged", "c0_readme", "c4_nn_cmf_r8"])
def test_many_rank_run_matches_reference_trajectory(name, world, tmp_path):
    """4 and 8 ranks (what SCALE runs): with 6-15 slices some ranks hold one slice and, at 8 ranks on the 6-slice case,
    some hold NONE — the empty-shard paths (zero-row kernels, stale reduction buffers) are part of the contract.
    c0_readme has constant_feasibility_penalty=True (MAX all-reduces) and a matrix-wise penalty on the sharded mode 0."""
    import torch.multiprocessing as mp

    g, X, R, kw = _load_case(name)
    k = min(20, g["A_traj"].shape[0])
    mp.spawn(_worker, args=(world, _free_port(), name, k, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for r in res[1:]:
        for key in ("A", "B", "C", "loss", "rec"):
            np.testing.assert_array_equal(res[0][key], r[key])
    r0 = res[0]
    errs = (_rel(r0["A"], g["A_traj"][k - 1]), _rel(r0["B"], g["B_traj"][k - 1]), _rel(r0["C"], g["C_traj"][k - 1]))
    print(f"[multi-rank parity] {name}, {world} ranks ({_backend(world)}): factor errors after {k} iterations {errs}")
    assert max(errs) < 1e-8, errs
    np.testing.assert_allclose(r0["loss"], g["regularized_loss"][: k + 1], rtol=1e-8)


def _worker_width(rank, world, port, name, k_iter, out_dir):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    from test_gpu_baseline_widths import twin

    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.distributed import make_shard

    dist = _init_group(rank, world, port)
    try:
        cfg, X = twin(name)
        sh = make_shard([x.shape[0] for x in X], rank, world, n_cols=X[0].shape[1])
        cmf, diag = cmf_aoadmm(X[sh.lo:sh.hi], cfg["R"], n_iter_max=k_iter, tol=None, absolute_tol=None,
                               return_errors=True, random_state=0, process_group=dist.group.WORLD, shard=sh,
                               **cfg["kw"])
        _, (A, B_is, C) = cmf
        if rank == 0:
            np.savez(os.path.join(out_dir, "rank0.npz"), A=A, B=np.concatenate(B_is, 0), C=C,
                     loss=np.asarray(diag.regularized_loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("c2", 4), ("c1", 8), ("c3", 2)])
def test_sharded_run_at_baseline_width(name, world, tmp_path):
    """The BASELINE-width twins (tests/test_gpu_baseline_widths.py: exact K / J-range / R / penalties of configs 1-3)
    sharded over 2 / 4 / 8 ranks against the live oracle after 20 outer iterations."""
    import torch.multiprocessing as mp

    sys.path.insert(0, HERE)
    from test_gpu_baseline_widths import oracle_trajectory, twin

    cfg, X = twin(name)
    traj, losses, _ = oracle_trajectory(name, X, cfg)
    k, A_o, B_o, C_o = next(t for t in traj if t[0] == 20)
    mp.spawn(_worker_width, args=(world, _free_port(), name, k, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    errs = (_rel(r0["A"], A_o), _rel(r0["B"], B_o), _rel(r0["C"], C_o))
    print(f"[multi-rank parity] {name} at BASELINE width, {world} ranks ({_backend(world)}): factor errors after "
          f"{k} iterations {errs}")
    assert max(errs) < 1e-8, errs
    np.testing.assert_allclose(r0["loss"], losses[: k + 1], rtol=1e-8)


def _mode0_case(variant=0, oracle=False):
    rs = np.random.RandomState(5)
    I, K, R = 23, 12, 3
    Js = rs.randint(6, 30, size=I)
    A = np.exp(-0.5 * ((np.arange(I)[:, None] - np.array([5.0, 11.0, 17.0])[None, :]) / 3.0) ** 2) + 0.05  # unimodal columns
    C = rs.uniform(size=(K, R))
    X = [(rs.uniform(size=(J, R)) * A[i]) @ C.T + 0.05 * rs.standard_normal(size=(J, K)) for i, J in enumerate(Js)]
    kw = dict(unimodal={0: True}, l2_norm_bound={0: 2.0}, non_negative=True, constant_feasibility_penalty=True,
              random_state=3)
    if variant == 1:  # TV + generalized L2 (graph Laplacian over the slice index) on mode 0: values enter the loss
        lap = 2 * np.eye(I) - np.eye(I, k=1) - np.eye(I, k=-1)
        lap[0, 0] = lap[-1, -1] = 1
        kw = dict(tv_penalty={0: 0.02}, generalized_l2_penalty={0: 0.1 * lap}, non_negative={1: True, 2: True},
                  constant_feasibility_penalty="A", random_state=3)
    elif variant == 2:  # unit simplex over all rows of A
        from matcouply_b200 import penalties as P
        from oracle import aoadmm_oracle as O

        cls = O.UnitSimplexP if oracle else P.UnitSimplex
        kw = dict(regs=[[cls()], [], []], non_negative={1: True, 2: True}, constant_feasibility_penalty=True,
                  random_state=3)
    return X, R, kw


def _worker_mode0(rank, world, port, k_iter, out_dir, variant=0):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.distributed import make_shard

    dist = _init_group(rank, world, port)
    try:
        X, R, kw = _mode0_case(variant)
        sh = make_shard([x.shape[0] for x in X], rank, world, n_cols=X[0].shape[1])
        cmf, diag = cmf_aoadmm(X[sh.lo:sh.hi], R, n_iter_max=k_iter, tol=None, absolute_tol=None, return_errors=True,
                               process_group=dist.group.WORLD, shard=sh, **kw)
        _, (A, B_is, C) = cmf
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), A=A, B=np.concatenate(B_is, 0), C=C,
                 loss=np.asarray(diag.regularized_loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_two_rank_matrix_penalties_on_sharded_mode0(tmp_path, variant):
    """Matrix-wise penalties on mode 0 (whole columns of A, whose rows are sharded): all-reduced column norms (L2Ball),
    gathered prox (Unimodality, TV, generalized L2, unit simplex) and the loss values of TV / generalized L2 evaluated on
    the gathered A must reproduce the single-process reference algorithm (oracle)."""
    import torch.multiprocessing as mp

    from oracle import aoadmm_oracle as O

    k = 15
    X, R, kw = _mode0_case(variant, oracle=True)
    o = O.ao_admm(X, R, n_iter_max=k, tol=None, absolute_tol=None, **kw)
    mp.spawn(_worker_mode0, args=(2, _free_port(), k, str(tmp_path), variant), nprocs=2, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    for key in ("A", "B", "C", "loss"):
        np.testing.assert_array_equal(r0[key], r1[key])
    errs = (_rel(r0["A"], o["A"]), _rel(r0["B"], np.concatenate(o["B_is"], 0)), _rel(r0["C"], o["C"]))
    assert max(errs) < 1e-8, errs
    np.testing.assert_allclose(r0["loss"], o["regularized_loss"][: k + 1], rtol=1e-8)


def _jump_worker(rank, port, name, out_path):
    """One GPU, a one-rank NCCL group: play rank 0 and rank 1 of a two-rank sharding one after the other (the
    all-reduces are identities, so the numbers are not the global solution — but the INITIAL STATE path is the sharded
    one) with the global stream walk (B2_MT_JUMP=0) and with the jump-ahead draw of the local rows (B2_MT_JUMP=1)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from matcouply_b200 import cmf_aoadmm
    from matcouply_b200.distributed import make_shard

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))

    def flat(v, out):
        if isinstance(v, np.ndarray):
            out.append(np.array(v))
        elif isinstance(v, (list, tuple)) or hasattr(v, "__iter__"):
            for u in v:
                flat(u, out)
        return out

    try:
        g, X, R, kw = _load_case(name)
        kw = dict(kw)
        kw.pop("n_iter_max", None)
        worst, n_arrays = 0.0, 0
        for fake_rank in (0, 1):
            sh = make_shard([x.shape[0] for x in X], fake_rank, 2)
            runs = {}
            for mode in ("0", "1"):
                os.environ["B2_MT_JUMP"] = mode
                cmf, admm, diag = cmf_aoadmm(X[sh.lo:sh.hi], R, n_iter_max=2, tol=None, absolute_tol=None,
                                             return_errors=True, return_admm_vars=True,
                                             process_group=dist.group.WORLD, shard=sh, **kw)
                arrays = flat([cmf[1][0], list(cmf[1][1]), cmf[1][2]], [])
                flat([list(admm[0]), list(admm[1])], arrays)
                arrays.append(np.asarray(diag.regularized_loss))
                runs[mode] = arrays
            assert len(runs["0"]) == len(runs["1"])
            for a, b in zip(runs["0"], runs["1"]):
                assert a.shape == b.shape
                worst = max(worst, float(np.max(np.abs(a - b))) if a.size else 0.0)
                n_arrays += 1
        with open(out_path, "w") as f:
            json.dump({"worst": worst, "arrays": n_arrays}, f)
    finally:
        os.environ.pop("B2_MT_JUMP", None)
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c2_nn_pf2_l1_ragged", "c3_unimodal_l2ball_pf2"])
def test_sharded_jump_ahead_draw_is_bit_identical_to_the_global_walk(name, tmp_path):
    """Runs on ONE GPU.  The sharded initial state drawn with the MT19937 jump-ahead (local rows only) leads to
    bit-identical factors, ADMM variables and losses as walking the whole global RandomState stream."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "jump.json")
    mp.spawn(_jump_worker, args=(_free_port(), name, out), nprocs=1, join=True)
    res = json.load(open(out))
    assert res["arrays"] > 10 and res["worst"] == 0.0, res


def test_device_keyword_binds_every_launch_to_that_device():
    """`cmf_aoadmm(..., device="cuda:1")` while device 0 is current: streams, workspaces, events and kernel launches
    must all bind to device 1 (the C ABI launches on torch's current stream).  Needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from matcouply_b200 import PackedMatrices, cmf_aoadmm

    g, X, R, kw = _load_case("c2_nn_pf2_l1_ragged")
    kw = dict(kw, n_iter_max=8, tol=None, absolute_tol=None)
    torch.cuda.set_device(0)
    ref = cmf_aoadmm(X, R, **kw)
    got = cmf_aoadmm(X, R, device="cuda:1", **kw)
    assert torch.cuda.current_device() == 0
    for a, b in zip((ref[1][0], np.concatenate(ref[1][1]), ref[1][2]), (got[1][0], np.concatenate(got[1][1]), got[1][2])):
        np.testing.assert_array_equal(a, b)
    packed = PackedMatrices.from_list(X, torch.float64, torch.device("cuda", 1))
    got2 = cmf_aoadmm(packed, R, **kw)  # device-resident input: the fit follows the data
    np.testing.assert_array_equal(got2[1][2], ref[1][2])
    with pytest.raises(ValueError):
        cmf_aoadmm(packed, R, device="cuda:0", **kw)
