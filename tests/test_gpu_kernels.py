"""Operator-level parity of every CUDA kernel (called through the C ABI) against NumPy float64 / the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _imports():
    from matcouply_b200 import _lib, _ops
    from oracle import aoadmm_oracle as O

    return _lib, _ops, O


def dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype or torch.float64).cuda()


def packed_x(N, K, dtype, rs):
    from matcouply_b200 import _ops

    ld = _ops.padded_ld(K, dtype)
    Xh = rs.standard_normal(size=(N, K))
    X = torch.zeros((N, ld), dtype=dtype, device="cuda")
    X[:, :K] = dev(Xh, dtype)
    return Xh if dtype == torch.float64 else Xh.astype(np.float32).astype(np.float64), X


XSTREAM_SHAPES = [(1, 1, 1), (37, 15, 3), (128, 32, 4), (300, 33, 5), (1000, 512, 16), (2500, 100, 20), (777, 70, 32),
                  (513, 1030, 8), (5000, 256, 12),
                  # R = 8 b + 1..4: tensor-core blocks + DFMA remainder columns (EX = 2 / 4)
                  (400, 64, 9), (401, 64, 10), (333, 40, 18), (500, 96, 26), (450, 72, 28), (300, 48, 11), (299, 80, 17)]


@pytest.mark.parametrize("N,K,R", XSTREAM_SHAPES)
@pytest.mark.parametrize("variant", ["fma", "dmma"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_xstream_y(N, K, R, variant, dtype):
    _lib, _ops, _ = _imports()
    if dtype == "f32" and variant == "dmma":
        pytest.skip("DMMA is fp64 only")
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(N * 7 + K)
    Xh, X = packed_x(N, K, tdt, rs)
    C = rs.standard_normal(size=(K, R))
    Cd = dev(C, tdt)
    Y = torch.full((N, R), float("nan"), dtype=tdt, device="cuda")
    ws = _ops.Workspace("cuda", K, R, tdt)
    _ops.xstream_y(X, N, K, Cd, Y, ws, _lib.VARIANT_FMA if variant == "fma" else _lib.VARIANT_DMMA)
    ref = Xh @ Cd.double().cpu().numpy()
    got = Y.double().cpu().numpy()
    scale = np.abs(Xh) @ np.abs(C) + 1e-300
    tol = 1e-13 if dtype == "f64" else 2e-5
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / scale) < tol, np.max(np.abs(got - ref) / scale)


@pytest.mark.parametrize("N,K,R", XSTREAM_SHAPES + [(513, 1024, 8), (64, 1030, 16), (700, 2048, 32), (3000, 1024, 20),
                                                    (20000, 256, 8), (9000, 128, 7)])  # narrow K: several CTAs per SM
@pytest.mark.parametrize("variant", ["fma", "dmma"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_xstream_z(N, K, R, variant, dtype):
    _lib, _ops, _ = _imports()
    if dtype == "f32" and variant == "dmma":
        pytest.skip("DMMA is fp64 only")
    tdt = torch.float64 if dtype == "f64" else torch.float32
    var = _lib.VARIANT_FMA if variant == "fma" else _lib.VARIANT_DMMA
    rs = np.random.RandomState(N * 11 + K)
    Xh, X = packed_x(N, K, tdt, rs)
    W = rs.standard_normal(size=(N, R))
    Wpad = _ops.alloc_w(N, R, tdt, "cuda", var)
    Wpad[:N, :R] = dev(W, tdt)
    Z = torch.full((K, R), float("nan"), dtype=tdt, device="cuda")
    ws = _ops.Workspace("cuda", K, R, tdt)
    _ops.xstream_z(X, N, K, Wpad, Z, ws, var)
    Wh = Wpad[:N, :R].double().cpu().numpy()
    ref = Xh.T @ Wh
    got = Z.double().cpu().numpy()
    scale = np.abs(Xh).T @ np.abs(Wh) + 1e-300
    tol = 1e-13 if dtype == "f64" else 5e-5
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref) / scale) < tol, np.max(np.abs(got - ref) / scale)


def test_xstream_linearity_large():
    """Size-independent property at a streaming size (X larger than L2): Y(X, C1 + C2) = Y(X, C1) + Y(X, C2) and
    Z/Y adjointness <X C, W> = <C, X^T W>."""
    _lib, _ops, _ = _imports()
    N, K, R = 200_000, 512, 16  # 0.8 GB
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    C1 = torch.randn((K, R), dtype=torch.float64, device="cuda", generator=g)
    C2 = torch.randn((K, R), dtype=torch.float64, device="cuda", generator=g)
    W = torch.randn((N, R), dtype=torch.float64, device="cuda", generator=g)
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    Zs = []
    for variant in (_lib.VARIANT_FMA, _lib.VARIANT_DMMA):
        Wp = _ops.alloc_w(N, R, torch.float64, "cuda", variant)
        Wp[:N, :R] = W
        Zv = torch.empty((K, R), dtype=torch.float64, device="cuda")
        _ops.xstream_z(X, N, K, Wp, Zv, ws, variant)
        Zs.append(Zv)
    assert torch.max(torch.abs(Zs[0] - Zs[1])).item() < 1e-8
    Z = Zs[1]
    Ys = []
    for variant in (_lib.VARIANT_FMA, _lib.VARIANT_DMMA):
        Y1, Y2, Y12 = (torch.empty((N, R), dtype=torch.float64, device="cuda") for _ in range(3))
        _ops.xstream_y(X, N, K, C1, Y1, ws, variant)
        _ops.xstream_y(X, N, K, C2, Y2, ws, variant)
        _ops.xstream_y(X, N, K, (C1 + C2).contiguous(), Y12, ws, variant)
        assert torch.max(torch.abs(Y12 - (Y1 + Y2))).item() < 1e-10
        Ys.append(Y1)
    assert torch.max(torch.abs(Ys[0] - Ys[1])).item() < 1e-10
    lhs = torch.sum(Ys[0] * W).item()
    rhs = torch.sum(C1 * Z).item()
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _ops.sumsq(X, N, K, out, ws)
    assert abs(out.item() - torch.sum(X * X).item()) < 1e-9 * out.item()


@pytest.mark.parametrize("n,R", [(1, 1), (50, 3), (1000, 16), (4097, 20), (300, 32)])
def test_gram_and_scale_gram(n, R):
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(n + R)
    M = rs.standard_normal(size=(n, R))
    G = torch.empty((R, R), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", 8, R, torch.float64)
    _ops.gram(dev(M), n, G, ws)
    np.testing.assert_allclose(G.cpu().numpy(), M.T @ M, rtol=1e-12, atol=1e-12 * n)
    A = rs.uniform(size=(7, R))
    lhs = torch.empty((7, R, R), dtype=torch.float64, device="cuda")
    _ops.scale_gram(G, dev(A), lhs)
    Gh = G.cpu().numpy()
    ref = np.stack([np.transpose(np.transpose(Gh * a) * a) for a in A])
    np.testing.assert_allclose(lhs.cpu().numpy(), ref, rtol=1e-14)
    rho = torch.empty(7, dtype=torch.float64, device="cuda")
    rmax = torch.empty(1, dtype=torch.float64, device="cuda")
    _ops.rho_from_trace(lhs, 7, R, 1.7, rho, rmax)
    ref_rho = np.array([0.5 * np.trace(L) * 1.7 for L in ref])
    np.testing.assert_allclose(rho.cpu().numpy(), ref_rho, rtol=1e-13)
    assert abs(rmax.item() - ref_rho.max()) < 1e-13 * ref_rho.max()


@pytest.mark.parametrize("R", [1, 2, 3, 8, 16, 20, 32])
@pytest.mark.parametrize("constant", [False, True])
def test_factor_batch(R, constant):
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(R)
    G = 9
    Ms = rs.standard_normal(size=(G, 3 * R + 2, R))
    lhs = np.stack([m.T @ m for m in Ms])
    rho = np.array([0.5 * np.trace(L) for L in lhs])
    rho_d, rmax = dev(rho), dev(np.array([rho.max()]))
    Minv = torch.empty((G, R, R), dtype=torch.float64, device="cuda")
    _ops.factor_batch(dev(lhs), G, R, rho_d, rmax if constant else None, 2, 0.3, Minv)
    eff = np.full(G, rho.max()) if constant else rho
    np.testing.assert_allclose(rho_d.cpu().numpy(), eff, rtol=1e-15)
    for g in range(G):
        full = lhs[g] + np.eye(R) * (eff[g] * 2 + 0.3)
        U, s, Uh = np.linalg.svd(full)  # the reference applies x (U/s) Uh  (decomposition.py:172,194)
        ref = (U / s) @ Uh
        np.testing.assert_allclose(Minv[g].cpu().numpy(), ref, rtol=1e-11, atol=1e-13 * np.abs(ref).max())


@pytest.mark.parametrize("R", [2, 5, 16, 20, 32])
def test_factor_batch_not_positive_definite_takes_the_eigen_route(R):
    """A symmetric lhs that is NOT positive definite (negative l2_penalty, or round-off on a rank-deficient Gram): the
    reference's `x (U / s) Uh` (decomposition.py:172, 256, 321) is the plain inverse for ANY invertible symmetric
    matrix.  The Cholesky pivot test must route such matrices to the Jacobi eigen-inverse (no silent NaN), leave the
    positive definite ones of the same batch on the Cholesky path, and agree with the reference's SVD route."""
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(100 + R)
    G = 11
    lhs = []
    for g in range(G):
        Q, _ = np.linalg.qr(rs.standard_normal(size=(R, R)))
        lam = rs.uniform(0.5, 3.0, size=R)
        if g % 2 == 0:  # indefinite, well conditioned
            lam[rs.randint(R)] *= -1.0
        lhs.append((Q * lam) @ Q.T)
    lhs = np.stack(lhs)
    lhs = 0.5 * (lhs + lhs.transpose(0, 2, 1))
    rho_d = dev(np.zeros(G))
    Minv = torch.empty((G, R, R), dtype=torch.float64, device="cuda")
    _ops.factor_batch(dev(lhs), G, R, rho_d, None, 0, 0.0, Minv)
    got = Minv.cpu().numpy()
    assert np.all(np.isfinite(got))
    for g in range(G):
        U, s, Uh = np.linalg.svd(lhs[g])
        ref = (U / s) @ Uh
        np.testing.assert_allclose(got[g], ref, rtol=1e-9, atol=1e-11 * np.abs(ref).max())


def ragged(rs, G, lo, hi, R):
    sizes = rs.randint(lo, hi + 1, size=G)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return sizes, off, rs.standard_normal(size=(int(off[-1]), R))


@pytest.mark.parametrize("R", [3, 16, 20, 32])
def test_slice_cross_and_rowscale(R):
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(R)
    G = 13
    sizes, off, B = ragged(rs, G, 1, 150, R)
    Y = rs.standard_normal(size=B.shape)
    CtC = rs.standard_normal(size=(R, R))
    cross = torch.empty((G, R, R), dtype=torch.float64, device="cuda")
    rhs = torch.empty((G, R), dtype=torch.float64, device="cuda")
    _ops.slice_cross(dev(B), dev(Y), dev(off, torch.int64), G, R, dev(CtC), cross, rhs)
    for g in range(G):
        Bg, Yg = B[off[g]:off[g + 1]], Y[off[g]:off[g + 1]]
        np.testing.assert_allclose(cross[g].cpu().numpy(), (Bg.T @ Bg) * CtC, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(rhs[g].cpu().numpy(), np.diag(Bg.T @ Yg), rtol=1e-12, atol=1e-12)
    A = rs.uniform(size=(G, R))
    gor = np.repeat(np.arange(G), sizes).astype(np.int32)
    W = torch.zeros((B.shape[0], R + 4), dtype=torch.float64, device="cuda")
    _ops.rowscale(dev(B), dev(A), dev(gor, torch.int32), B.shape[0], R, W)
    np.testing.assert_array_equal(W[:, :R].cpu().numpy(), B * A[gor])
    assert float(W[:, R:].abs().max()) == 0.0
    G = torch.empty((R, R), dtype=torch.float64, device="cuda")
    _ops.gram(W, B.shape[0], G, _ops.Workspace("cuda", 8, R, torch.float64))
    np.testing.assert_allclose(G.cpu().numpy(), (B * A[gor]).T @ (B * A[gor]), rtol=1e-12, atol=1e-12)


def _penalty_cases(O):
    return [
        ("nonneg", O.NonNeg(), 0, False, 0.0, 0.0),
        ("box", O.BoxP(-0.3, 0.6), 1, False, -0.3, 0.6),
        ("l1", O.L1P(0.25), 2, False, 0.25, 0.0),
        ("l1nn", O.L1P(0.25, non_negativity=True), 2, True, 0.25, 0.0),
    ]


@pytest.mark.parametrize("mode", ["single", "indexed", "identity"])
@pytest.mark.parametrize("R", [3, 16, 20])
def test_admm_solve_elementwise(mode, R):
    """x-update + prox + dual update for all elementwise penalties at once vs the reference formulas."""
    _lib, _ops, O = _imports()
    rs = np.random.RandomState(R + len(mode))
    G = 6
    sizes, off, rhs = ragged(rs, G, 2, 40, R)
    n = rhs.shape[0]
    if mode == "single":
        gor, ngroups, gmode = np.zeros(n, dtype=np.int32), 1, _lib.GROUP_SINGLE
    elif mode == "indexed":
        gor, ngroups, gmode = np.repeat(np.arange(G), sizes).astype(np.int32), G, _lib.GROUP_INDEXED
    else:
        gor, ngroups, gmode = np.arange(n, dtype=np.int32), n, _lib.GROUP_IDENTITY
    rho = rs.uniform(0.5, 2.0, size=ngroups)
    Ms = rs.standard_normal(size=(ngroups, 2 * R, R))
    Minv = np.stack([np.linalg.inv(m.T @ m + np.eye(R)) for m in Ms])
    scale = rs.uniform(0.5, 1.5, size=(ngroups, R))
    cases = _penalty_cases(O)
    aux = [rs.standard_normal(size=(n, R)) for _ in cases]
    dual = [rs.standard_normal(size=(n, R)) for _ in cases]
    aux_d, dual_d = [dev(a) for a in aux], [dev(d) for d in dual]
    descs = _ops.make_descs([(c[2], c[3], c[4], c[5], a, d) for c, a, d in zip(cases, aux_d, dual_d)])
    x = torch.empty((n, R), dtype=torch.float64, device="cuda")
    _ops.admm_solve(n, R, dev(rhs), dev(scale), gmode, dev(gor, torch.int32), dev(rho), dev(Minv), descs, len(cases), x)
    # reference
    shifted = sum(a - d for a, d in zip(aux, dual))
    s = rho[gor][:, None] * shifted + rhs * scale[gor]
    x_ref = np.einsum("nr,nrc->nc", s, Minv[gor])
    np.testing.assert_allclose(x.cpu().numpy(), x_ref, rtol=1e-11, atol=1e-12)
    for (name, pen, *_), a_d, d_d, d0 in zip(cases, aux_d, dual_d, dual):
        v = x_ref + d0
        z_ref = np.stack([pen.prox(v[i], rho[gor[i]], None) for i in range(n)])
        np.testing.assert_allclose(a_d.cpu().numpy(), z_ref, rtol=1e-10, atol=1e-11, err_msg=name)
        np.testing.assert_allclose(d_d.cpu().numpy(), v - z_ref, rtol=1e-10, atol=1e-11, err_msg=name)


@pytest.mark.parametrize("R", [1, 3, 8, 16, 20, 32])
@pytest.mark.parametrize("n_pen", [0, 1, 2])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("grouped", [False, True, "mma"])
def test_admm_local_equals_iterated_admm_solve(R, n_pen, dtype, grouped):
    """The fused whole-inner-loop kernel must reproduce 5 launches of the one-iteration kernel (and W = x o a).
    grouped=True: CTA-per-slice shuffle kernel, "mma": its tensor-core (DMMA) formulation."""
    _lib, _ops, O = _imports()
    _lib.load().b2_set_option(_lib.OPT_ADMM_LOCAL_MMA, 1 if grouped == "mma" else 0)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(R * 10 + n_pen)
    G = 7
    sizes, off, rhs = ragged(rs, G, 1, 50, R)
    n = rhs.shape[0]
    gor = dev(np.repeat(np.arange(G), sizes).astype(np.int32), torch.int32)
    rho = dev(rs.uniform(0.5, 2.0, size=G), tdt)
    Ms = rs.standard_normal(size=(G, 2 * R, R))
    Minv = dev(np.stack([np.linalg.inv(m.T @ m + np.eye(R)) for m in Ms]), tdt)
    scale = dev(rs.uniform(0.5, 1.5, size=(G, R)), tdt)
    cases = [(_lib.PEN_L1, False, 0.2, 0.0), (_lib.PEN_BOX, False, -0.1, 0.8)][:n_pen]
    aux0 = [rs.standard_normal(size=(n, R)) for _ in cases]
    dual0 = [rs.standard_normal(size=(n, R)) for _ in cases]
    results = []
    for fused in (False, True):
        aux, dual = [dev(a, tdt) for a in aux0], [dev(d, tdt) for d in dual0]
        descs = _ops.make_descs([(c[0], c[1], c[2], c[3], a, d) for c, a, d in zip(cases, aux, dual)])
        x = torch.zeros((n, R), dtype=tdt, device="cuda")
        W = torch.zeros((n, R + 4), dtype=tdt, device="cuda")
        BtB = torch.zeros((G, R, R), dtype=tdt, device="cuda")
        if fused and grouped:  # CTA-per-slice path: Minv staged in shared memory, B^T B fused
            _ops.admm_local(n, R, dev(rhs, tdt), scale, _lib.GROUP_INDEXED, gor, rho, Minv, descs, n_pen, 5, x, W,
                            row_off=dev(off, torch.int64), n_groups=G, BtB_out=BtB)
            xh = x.double().cpu().numpy()
            ref = np.stack([xh[off[g]:off[g + 1]].T @ xh[off[g]:off[g + 1]] for g in range(G)])
            np.testing.assert_allclose(BtB.double().cpu().numpy(), ref, rtol=1e-11 if dtype == "f64" else 1e-4,
                                       atol=1e-11 if dtype == "f64" else 1e-3)
        elif fused:
            _ops.admm_local(n, R, dev(rhs, tdt), scale, _lib.GROUP_INDEXED, gor, rho, Minv, descs, n_pen, 5, x, W)
        else:
            for _ in range(5):
                _ops.admm_solve(n, R, dev(rhs, tdt), scale, _lib.GROUP_INDEXED, gor, rho, Minv, descs, n_pen, x)
            _ops.rowscale(x, scale, gor, n, R, W)
        results.append([t.double().cpu().numpy() for t in [x, W] + aux + dual])
    _lib.load().b2_set_option(_lib.OPT_ADMM_LOCAL_MMA, 1)
    tol = 1e-11 if dtype == "f64" else 2e-4
    for a, b in zip(*results):
        np.testing.assert_allclose(a, b, rtol=tol, atol=tol)


def test_prox_golden_vectors(golden_dir):
    """Stand-alone penalty objects (the reference's ADMMPenalty protocol) against vectors produced by the reference."""
    from matcouply_b200 import penalties as P

    g = np.load(os.path.join(golden_dir, "operators.npz"))
    M = g["prox_in"]
    np.testing.assert_array_equal(P.NonNegativity().factor_matrix_update(M, 1.3, None), g["prox_nonneg"])
    np.testing.assert_array_equal(P.Box(-0.5, 0.7).factor_matrix_update(M, 1.3, None), g["prox_box"])
    np.testing.assert_allclose(P.L1Penalty(0.4).factor_matrix_update(M, 1.3, None), g["prox_l1"], rtol=1e-15,
                               atol=1e-16)
    np.testing.assert_allclose(P.L1Penalty(0.4, non_negativity=True).factor_matrix_update(M, 1.3, None),
                               g["prox_l1_nn"], rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(P.L2Ball(1.5).factor_matrix_update(M, 1.3, None), g["prox_l2ball"], rtol=1e-14)
    np.testing.assert_allclose(P.L2Ball(1.5, non_negativity=True).factor_matrix_update(M, 1.3, None),
                               g["prox_l2ball_nn"], rtol=1e-14)
    row = P.L1Penalty(0.4).factor_matrix_row_update(M[3], 1.3, None)
    np.testing.assert_allclose(row, g["prox_l1"][3], rtol=1e-15, atol=1e-16)
    assert abs(P.L1Penalty(0.4).penalty(M) - 0.4 * np.abs(M).sum()) < 1e-12
    assert P.NonNegativity().penalty(M) == 0


def test_l2ball_groups():
    _lib, _ops, O = _imports()
    rs = np.random.RandomState(3)
    for R in (1, 5, 16, 32):
        for nn in (False, True):
            sizes, off, V = ragged(rs, 9, 1, 300, R)
            V *= 3
            aux = torch.empty(V.shape, dtype=torch.float64, device="cuda")
            dual = dev(V)
            _ops.prox_l2ball(aux, dual, dev(off, torch.int64), 9, R, 1.25, nn)
            pen = O.L2BallP(1.25, non_negativity=nn)
            ref = np.concatenate([pen.prox(V[off[g]:off[g + 1]], 1.0, None) for g in range(9)], 0)
            np.testing.assert_allclose(aux.cpu().numpy(), ref, rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(dual.cpu().numpy(), V - ref, rtol=1e-12, atol=1e-14)


@pytest.fixture(params=[None, 1, 6, 9, 10, 11, 12, 15],
                ids=["variant_default", "variant_1", "variant_6", "variant_9", "variant_10", "variant_11", "variant_12",
                     "variant_15"])
def unimodal_variant(request):
    """Runs a unimodal test under the default kernel variant, the IEEE-division one (1), the reciprocal-division /
    256-bit record one (6), merge + finalisation in one trip (9), the compact-prefix-error variants (10-12) and the deferred-fill ones (14 = default, 15):
    bit-exactness must not depend on the variant."""
    from matcouply_b200 import _lib

    lib = _lib.load()
    before = lib.b2_get_option(_lib.OPT_UNIMODAL_VARIANT)
    if request.param is not None:
        lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, request.param)
    yield request.param
    lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, before)


def test_exact_division_by_counts_selftest():
    """The division of the unimodal kernel's fast variants (reciprocal + two Markstein corrections) equals the IEEE
    division bit for bit: 2^28 pseudo-random numerators (all exponents, zeros, subnormals, near-multiples) x divisors
    1..4096 on the device."""
    _lib, _ops, O = _imports()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed, max_cnt in ((1, 4096), (2, 17), (3, 1 << 20)):
        _lib.call("b2_selftest_div_count", 1 << 28, seed, max_cnt, bad.data_ptr(), _ops._stream())
        assert int(bad.item()) == 0, (seed, max_cnt, int(bad.item()))


def test_unimodal_golden_bit_exact(golden_dir, unimodal_variant):
    """Unimodal regression: fits AND peak indices bit-identical to the reference on its own golden vectors."""
    _lib, _ops, O = _imports()
    g = np.load(os.path.join(golden_dir, "operators.npz"))
    ws = _ops.Workspace("cuda", 1, 4, torch.float64)
    for k in range(int(g["uni_count"])):
        y, fit, peaks, nn = g[f"uni{k}_y"], g[f"uni{k}_fit"], g[f"uni{k}_peaks"], bool(g[f"uni{k}_nn"])
        n, R = y.shape
        aux = torch.empty((n, R), dtype=torch.float64, device="cuda")
        dual = dev(y)
        pk = torch.zeros(R, dtype=torch.int32, device="cuda")
        _ops.prox_unimodal(aux, dual, dev(np.array([0, n]), torch.int64), 1, R, n, nn, ws, pk)
        assert np.array_equal(pk.cpu().numpy(), peaks), (k, pk.cpu().numpy(), peaks)
        assert np.array_equal(aux.cpu().numpy(), fit), (k, np.abs(aux.cpu().numpy() - fit).max())
        np.testing.assert_array_equal(dual.cpu().numpy(), y - fit)


def test_unimodal_ragged_groups_vs_oracle_and_sklearn(unimodal_variant):
    _lib, _ops, O = _imports()
    from sklearn.isotonic import IsotonicRegression

    rs = np.random.RandomState(11)
    R, G = 6, 40
    sizes, off, V = ragged(rs, G, 1, 260, R)
    ws = _ops.Workspace("cuda", 1, R, torch.float64)
    for nn in (False, True):
        aux = torch.empty(V.shape, dtype=torch.float64, device="cuda")
        dual = dev(V)
        pk = torch.zeros(G * R, dtype=torch.int32, device="cuda")
        _ops.prox_unimodal(aux, dual, dev(off, torch.int64), G, R, int(sizes.max()), nn, ws, pk)
        got, pk = aux.cpu().numpy(), pk.cpu().numpy().reshape(G, R)
        for g in range(G):
            ref, peaks = O.unimodal_regression(V[off[g]:off[g + 1]], nn, return_peaks=True)
            assert np.array_equal(pk[g], peaks)
            assert np.array_equal(got[off[g]:off[g + 1]], ref)
        # independent check of optimality on one column: best split of two isotonic fits (sklearn)
        g = int(np.argmax(sizes))
        y = V[off[g]:off[g + 1], 0]
        n = len(y)
        best = np.inf
        for t in range(n + 1):
            left = IsotonicRegression(y_min=0 if nn else None).fit_transform(np.arange(t), y[:t]) if t else np.zeros(0)
            right = (IsotonicRegression(increasing=False, y_min=0 if nn else None).fit_transform(np.arange(n - t), y[t:])
                     if n - t else np.zeros(0))
            best = min(best, np.sum((np.concatenate([left, right]) - y) ** 2))
        assert abs(np.sum((got[off[g]:off[g + 1], 0] - y) ** 2) - best) < 1e-9 * max(best, 1.0)


def test_unimodal_few_long_groups_all_variants_bit_identical():
    """Few long ragged groups (the deferred fill splits the rows of a slice over several CTAs; block counts and prefix
    lengths that are not multiples of four): every kernel variant gives the oracle's fit and peaks bit for bit."""
    _lib, _ops, O = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(5)
    R = 5
    sizes = np.array([5003, 1, 2998, 7, 4])
    off = np.concatenate([[0], np.cumsum(sizes)])
    t = np.arange(off[-1])[:, None]
    V = np.exp(-0.5 * ((t % 3000 - 1400.0) / 300.0) ** 2) + 0.05 * rs.standard_normal(size=(off[-1], R))
    before = lib.b2_get_option(_lib.OPT_UNIMODAL_VARIANT)
    try:
        for nn in (False, True):
            refs = [O.unimodal_regression(V[off[g]:off[g + 1]], nn, return_peaks=True) for g in range(len(sizes))]
            fit = np.concatenate([r[0] for r in refs], 0)
            peaks = np.concatenate([r[1] for r in refs])
            for variant in (9, 10, 11, 13, 14, 15):
                lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, variant)
                ws = _ops.Workspace("cuda", 1, R, torch.float64)
                aux = torch.empty(V.shape, dtype=torch.float64, device="cuda")
                dual = dev(V)
                pk = torch.zeros(len(sizes) * R, dtype=torch.int32, device="cuda")
                _ops.prox_unimodal(aux, dual, dev(off, torch.int64), len(sizes), R, int(sizes.max()), nn, ws, pk)
                assert np.array_equal(pk.cpu().numpy(), peaks), variant
                assert np.array_equal(aux.cpu().numpy(), fit), variant
                np.testing.assert_array_equal(dual.cpu().numpy(), V - fit)
    finally:
        lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, before)


def test_unimodal_more_columns_than_scratch_slots():
    """More columns than scratch slots (148 SMs x 1024): the kernel works in rounds and the deferred-fill variants fall
    back to the fill inside the PAVA kernel; same bits as the reference order of operations (oracle on a sample)."""
    _lib, _ops, O = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(8)
    R, G = 8, 148 * 1024 // 8 + 300
    sizes, off, V = ragged(rs, G, 3, 12, R)
    before = lib.b2_get_option(_lib.OPT_UNIMODAL_VARIANT)
    outs = {}
    try:
        for variant in (9, 14):
            lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, variant)
            ws = _ops.Workspace("cuda", 1, R, torch.float64)
            aux = torch.empty(V.shape, dtype=torch.float64, device="cuda")
            dual = dev(V)
            pk = torch.zeros(G * R, dtype=torch.int32, device="cuda")
            _ops.prox_unimodal(aux, dual, dev(off, torch.int64), G, R, int(sizes.max()), True, ws, pk)
            outs[variant] = (aux.cpu().numpy(), dual.cpu().numpy(), pk.cpu().numpy())
    finally:
        lib.b2_set_option(_lib.OPT_UNIMODAL_VARIANT, before)
    for variant in (14,):
        for a, b in zip(outs[9], outs[variant]):
            assert np.array_equal(a, b), variant
    for g in (0, 1, G // 2, G - 2, G - 1):
        ref, peaks = O.unimodal_regression(V[off[g]:off[g + 1]], True, return_peaks=True)
        assert np.array_equal(outs[14][0][off[g]:off[g + 1]], ref)
        assert np.array_equal(outs[14][2][g * R:(g + 1) * R], peaks)


def test_parafac2_prox_golden(golden_dir):
    """Parafac2.factor_matrices_update (R x R Jacobi polar route) vs the reference's thin-SVD route."""
    from matcouply_b200 import penalties as P

    g = np.load(os.path.join(golden_dir, "operators.npz"))
    off = g["pf2_off"]
    fms = [g["pf2_in"][a:b] for a, b in zip(off[:-1], off[1:])]
    R = g["pf2_delta_in"].shape[0]
    bases, delta = P.Parafac2().factor_matrices_update(fms, list(g["pf2_rhos"]),
                                                       ([np.eye(f.shape[0], R) for f in fms], g["pf2_delta_in"]))
    np.testing.assert_allclose(np.concatenate(bases, 0), g["pf2_basis"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(delta, g["pf2_delta_out"], rtol=0, atol=1e-11)
    for b in bases:
        np.testing.assert_allclose(b.T @ b, np.eye(R), atol=1e-12)


@pytest.mark.parametrize("R", [2, 8, 20, 32])
def test_parafac2_polar_random(R):
    from matcouply_b200 import penalties as P
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(R)
    Js = [R, R + 1, 3 * R, 200, 57 + R]
    fms = [rs.standard_normal(size=(J, R)) for J in Js]
    rhos = list(rs.uniform(0.5, 2, size=len(Js)))
    delta = rs.uniform(size=(R, R)) + 0.5 * np.eye(R)
    bases, new_delta = P.Parafac2().factor_matrices_update(fms, rhos, (None, delta))
    ob, od = O.Parafac2P().prox_list(fms, rhos, (None, delta))
    for b, o in zip(bases, ob):
        np.testing.assert_allclose(b, o, atol=1e-9)
    np.testing.assert_allclose(new_delta, od, atol=1e-9)


def test_reduce_stats_and_fit_terms():
    _lib, _ops, O = _imports()
    rs = np.random.RandomState(0)
    n = 100_003
    x, y = rs.standard_normal(size=n), rs.standard_normal(size=n)
    ws = _ops.Workspace("cuda", 1, 4, torch.float64)
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    _ops.reduce_stats(dev(x), dev(y), n, out, ws)
    np.testing.assert_allclose(out.cpu().numpy(), [np.sum((x - y) ** 2), np.sum(x ** 2), np.sum(np.abs(x))], rtol=1e-12)
    G, R = 500, 7
    rhs, A = rs.standard_normal(size=(G, R)), rs.standard_normal(size=(G, R))
    cross = rs.standard_normal(size=(G, R, R))
    out2 = torch.zeros(2, dtype=torch.float64, device="cuda")
    _ops.fit_terms(dev(rhs), dev(cross), dev(A), G, R, out2, ws)
    ref = [np.sum(rhs * A), sum(a @ c @ a for a, c in zip(A, cross))]
    np.testing.assert_allclose(out2.cpu().numpy(), ref, rtol=1e-11)


def test_errors_are_reported_not_swallowed():
    _lib, _ops, _ = _imports()
    X = torch.zeros((4, 4), dtype=torch.float64, device="cuda")
    C = torch.zeros((4, 40), dtype=torch.float64, device="cuda")
    Y = torch.zeros((4, 40), dtype=torch.float64, device="cuda")
    ws = _ops.Workspace("cuda", 4, 32, torch.float64)
    with pytest.raises(RuntimeError, match="rank"):
        _ops.xstream_y(X, 4, 4, C, Y, ws)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        _ops.xstream_y(X.cpu(), 4, 4, C, Y, ws)


@pytest.mark.parametrize("R", [1, 3, 8, 12, 16, 20, 27, 32])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_slice_gram_family(R, dtype):
    """b2_slice_gram (DMMA), b2_slice_coldot, b2_weighted_gram_sum, b2_hadamard_bcast vs NumPy."""
    _lib, _ops, _ = _imports()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(R)
    G = 11
    sizes, off, B = ragged(rs, G, 1, 200, R)
    sizes[3] = 0  # an empty slice
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    B = B[: off[-1]]
    Y = rs.standard_normal(size=B.shape)
    A = rs.uniform(size=(G, R))
    CtC = rs.standard_normal(size=(R, R))
    Bd, Yd, offd = dev(B, tdt), dev(Y, tdt), dev(off, torch.int64)
    Bh, Yh = Bd.double().cpu().numpy(), Yd.double().cpu().numpy()
    BtB = torch.full((G, R, R), float("nan"), dtype=tdt, device="cuda")
    _ops.slice_gram(Bd, offd, G, R, BtB)
    ref = np.stack([Bh[off[g]:off[g + 1]].T @ Bh[off[g]:off[g + 1]] for g in range(G)])
    tol = 1e-12 if dtype == "f64" else 2e-5
    np.testing.assert_allclose(BtB.double().cpu().numpy(), ref, rtol=tol, atol=tol * 200)
    rhs = torch.full((G, R), float("nan"), dtype=tdt, device="cuda")
    _ops.slice_coldot(Bd, Yd, offd, G, R, rhs)
    ref_rhs = np.stack([np.sum(Bh[off[g]:off[g + 1]] * Yh[off[g]:off[g + 1]], axis=0) for g in range(G)])
    np.testing.assert_allclose(rhs.double().cpu().numpy(), ref_rhs, rtol=tol, atol=tol * 200)
    Ad = dev(A, tdt)
    Ah = Ad.double().cpu().numpy()
    lhs = torch.empty((R, R), dtype=tdt, device="cuda")
    _ops.weighted_gram_sum(BtB, Ad, G, R, lhs)
    BtBh = BtB.double().cpu().numpy()
    np.testing.assert_allclose(lhs.double().cpu().numpy(), sum(np.outer(a, a) * M for a, M in zip(Ah, BtBh)),
                               rtol=tol * 10, atol=tol * 200)
    cross = torch.empty((G, R, R), dtype=tdt, device="cuda")
    Cd = dev(CtC, tdt)
    _ops.hadamard_bcast(BtB, Cd, G, R, cross)
    np.testing.assert_allclose(cross.double().cpu().numpy(), BtBh * Cd.double().cpu().numpy(), rtol=tol)


@pytest.mark.parametrize("R", [2, 4, 6, 8, 12, 16, 20, 22, 28, 32, 5])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("mma", [1, 0])
def test_pf2_rowpass_both_formulations(R, dtype, mma):
    """b2_pf2_rowpass (DMMA tile kernel / shuffle kernel) vs a NumPy restatement of one fused inner iteration
    (decomposition.py:259-289, penalties.py:1256-1281): deferred and stored PARAFAC2 aux, an elementwise and a
    column-coupled companion penalty, ragged slices with an empty one, x / W / B^T B emission."""
    _lib, _ops, _ = _imports()
    lib = _lib.load()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    if dtype == "f32" and R % 4 != 0 and mma:
        pytest.skip("row size not a multiple of 16 bytes: the tensor-core kernel is not selected")
    rs = np.random.RandomState(100 + R)
    G = 7
    sizes, off, _ = ragged(rs, G, 1, 300, R)
    sizes[2] = 0
    sizes[5] = 64
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(off[-1])
    f = lambda *s: rs.standard_normal(size=s)  # noqa: E731
    Y, A, rho = f(N, R), rs.uniform(0.5, 1.5, size=(G, R)), rs.uniform(0.5, 2.0, size=G)
    Minv = np.stack([np.linalg.inv(m @ m.T + R * np.eye(R)) for m in f(G, R, R)])
    Wm, Delta = f(G, R, R) / np.sqrt(R), f(R, R)
    pf_aux, pf_dual, nn_aux, nn_dual, l2_aux, l2_dual = (f(N, R) for _ in range(6))
    tol = 1e-11 if dtype == "f64" else 3e-4
    lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, mma)
    try:
        for deferred in (0, 1):
            for last in (0, 1):
                d = {k: dev(v, tdt) for k, v in dict(Y=Y, A=A, rho=rho, Minv=Minv, Wm=Wm, Delta=Delta, pf_aux=pf_aux,
                                                     pf_dual=pf_dual, nn_aux=nn_aux, nn_dual=nn_dual, l2_aux=l2_aux,
                                                     l2_dual=l2_dual).items()}
                h = {k: v.double().cpu().numpy() for k, v in d.items()}  # the inputs as the kernel sees them
                descs = _ops.make_descs([(_lib.PEN_PARAFAC2, 0, 0, 0, d["pf_aux"], d["pf_dual"]),
                                         (_lib.PEN_NONNEG, 0, 0, 0, d["nn_aux"], d["nn_dual"]),
                                         (_lib.PEN_L2BALL, 1, 1.0, 0, d["l2_aux"], d["l2_dual"])])
                x = torch.full((N, R), float("nan"), dtype=tdt, device="cuda")
                Wp = _ops.alloc_w(N, R, tdt, "cuda")
                S = torch.full((G, R, R), float("nan"), dtype=tdt, device="cuda")
                BtB = torch.full((G, R, R), float("nan"), dtype=tdt, device="cuda")
                _ops.pf2_rowpass(dev(off, torch.int64), G, R, d["Y"], d["A"], d["rho"], d["Minv"], descs, 3, deferred,
                                 d["Wm"], d["Delta"], x if last else None, Wp if last else None, S,
                                 BtB if last else None)
                torch.cuda.synchronize()
                gor = np.repeat(np.arange(G), sizes)
                if deferred:
                    T = np.einsum("gik,kj->gij", h["Wm"], h["Delta"])
                    pd = np.einsum("nk,nkj->nj", h["pf_dual"], T[gor])
                    dpf = h["pf_dual"] - pd
                else:
                    pd, dpf = h["pf_aux"], h["pf_dual"]
                sh = (pd - dpf) + (h["nn_aux"] - h["nn_dual"]) + (h["l2_aux"] - h["l2_dual"])
                s = h["rho"][gor][:, None] * sh + h["Y"] * h["A"][gor]
                xr = np.einsum("nk,nkj->nj", s, h["Minv"][gor])
                vn = xr + dpf
                np.testing.assert_allclose(d["pf_dual"].double().cpu().numpy(), vn, rtol=tol, atol=tol * 10)
                v_nn = xr + h["nn_dual"]
                np.testing.assert_allclose(d["nn_aux"].double().cpu().numpy(), np.maximum(v_nn, 0), rtol=tol, atol=tol * 10)
                np.testing.assert_allclose(d["nn_dual"].double().cpu().numpy(), np.minimum(v_nn, 0), rtol=tol, atol=tol * 10)
                np.testing.assert_allclose(d["l2_dual"].double().cpu().numpy(), xr + h["l2_dual"], rtol=tol, atol=tol * 10)
                np.testing.assert_array_equal(d["l2_aux"].double().cpu().numpy(), h["l2_aux"])  # untouched here
                Sref = np.stack([vn[off[g]:off[g + 1]].T @ vn[off[g]:off[g + 1]] for g in range(G)])
                np.testing.assert_allclose(S.double().cpu().numpy(), Sref, rtol=tol * 10, atol=tol * 1e3)
                if last:
                    np.testing.assert_allclose(x.double().cpu().numpy(), xr, rtol=tol, atol=tol * 10)
                    np.testing.assert_allclose(Wp[:N, :R].double().cpu().numpy(), xr * h["A"][gor], rtol=tol, atol=tol * 10)
                    assert float(Wp[:N, R:].abs().max() if Wp.shape[1] > R else 0.0) == 0.0
                    Bref = np.stack([xr[off[g]:off[g + 1]].T @ xr[off[g]:off[g + 1]] for g in range(G)])
                    np.testing.assert_allclose(BtB.double().cpu().numpy(), Bref, rtol=tol * 10, atol=tol * 1e3)
    finally:
        lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, _lib.PF2_ROWPASS_DEFAULT)


@pytest.mark.parametrize("R", [4, 8, 12, 16, 20, 24, 28, 32])
@pytest.mark.parametrize("n_extra", [0, 1])
def test_pf2_rowpass_steady_state_kernel_bit_identical(R, n_extra):
    """The steady-state specialisation of the row pass (csrc/pf2_rowpass_v2.cuh: compile-time rank, deferred prox,
    PARAFAC2 alone or with non-negativity, full-tile fast path, last-releaser refill) performs the same operations in
    the same order as the general DMMA tile kernel: a whole B-update (first / middle / middle / last pass, the flag
    sequence of `_engine._step_B_pf2_fused`) must leave BIT-identical V, T / (aux, dual), x, W, S and B^T B.  Slices of
    0, 1, 63, 64, 65, 128 and ~1000 rows cover the empty, ragged-only, full-only and mixed tile paths and ring wrap."""
    _lib, _ops, _ = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(300 + R + n_extra)
    sizes = np.array([0, 1, 63, 64, 65, 128, 700 + 37 * (R % 5), 5, 1029, 64], dtype=np.int64)
    G = len(sizes)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(off[-1])
    f = lambda *s: rs.standard_normal(size=s)  # noqa: E731
    Y, A, rho = f(N, R), rs.uniform(0.5, 1.5, size=(G, R)), rs.uniform(0.5, 2.0, size=G)
    Minv = np.stack([np.linalg.inv(m @ m.T + R * np.eye(R)) for m in f(G, R, R)]) * 0.5
    Wm, Delta = f(G, R, R) / np.sqrt(R), f(R, R) / np.sqrt(R)
    init = dict(pf_aux=f(N, R), pf_dual=f(N, R), nn_aux=f(N, R), nn_dual=f(N, R))
    offd = dev(off, torch.int64)
    gor = np.repeat(np.arange(G), sizes)
    results = {}
    try:
        for opt in (1, 2):
            lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, opt)
            d = {k: dev(v) for k, v in init.items()}
            pens = [(_lib.PEN_PARAFAC2, 0, 0, 0, d["pf_aux"], d["pf_dual"])]
            if n_extra:
                pens.append((_lib.PEN_NONNEG, 0, 0, 0, d["nn_aux"], d["nn_dual"]))
            descs = _ops.make_descs(pens)
            x = torch.full((N, R), float("nan"), dtype=torch.float64, device="cuda")
            Wp = _ops.alloc_w(N, R, torch.float64, "cuda")
            BtB = torch.full((G, R, R), float("nan"), dtype=torch.float64, device="cuda")
            S_all = []
            n_pass = 4
            launches0 = int(lib.b2_launch_count())
            for it in range(n_pass):
                last = it == n_pass - 1
                flags = 1 | (2 if it > 0 else 0) | (0 if last else 4)
                S = torch.full((G, R, R), float("nan"), dtype=torch.float64, device="cuda")
                _ops.pf2_rowpass(offd, G, R, dev(Y), dev(A), dev(rho), dev(Minv), descs, len(pens), flags, dev(Wm),
                                 dev(Delta), x if last else None, Wp if last else None, S, BtB if last else None)
                S_all.append(S.cpu().numpy())
            torch.cuda.synchronize()
            assert int(lib.b2_launch_count()) - launches0 == n_pass  # one kernel per pass on either path
            results[opt] = dict(x=x.cpu().numpy(), W=Wp.cpu().numpy(), BtB=BtB.cpu().numpy(), S=np.stack(S_all),
                                **{k: v.cpu().numpy() for k, v in d.items()})
    finally:
        lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, _lib.PF2_ROWPASS_DEFAULT)
    for key in results[1]:
        if n_extra == 0 and key.startswith("nn_"):
            continue
        np.testing.assert_array_equal(results[2][key], results[1][key], err_msg=key)
    # and one pass against NumPy (deferred, first pass of a B-update)
    lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, 2)
    d = {k: dev(v) for k, v in init.items()}
    pens = [(_lib.PEN_PARAFAC2, 0, 0, 0, d["pf_aux"], d["pf_dual"])]
    if n_extra:
        pens.append((_lib.PEN_NONNEG, 0, 0, 0, d["nn_aux"], d["nn_dual"]))
    S = torch.full((G, R, R), float("nan"), dtype=torch.float64, device="cuda")
    _ops.pf2_rowpass(offd, G, R, dev(Y), dev(A), dev(rho), dev(Minv), _ops.make_descs(pens), len(pens), 1 | 4, dev(Wm),
                     dev(Delta), None, None, S, None)
    T = np.einsum("gik,kj->gij", Wm, Delta)
    pd = np.einsum("nk,nkj->nj", init["pf_dual"], T[gor])
    dpf = init["pf_dual"] - pd
    sh = pd - dpf
    if n_extra:
        sh = sh + init["nn_aux"] - init["nn_dual"]
    xr = np.einsum("nk,nkj->nj", rho[gor][:, None] * sh + Y * A[gor], Minv[gor])
    vn = xr + dpf
    np.testing.assert_allclose(d["pf_dual"].cpu().numpy(), vn, rtol=1e-11, atol=1e-10)
    if n_extra:
        np.testing.assert_allclose(d["nn_dual"].cpu().numpy(), xr + init["nn_dual"], rtol=1e-11, atol=1e-10)  # T' = x + dual
    Sref = np.stack([vn[off[g]:off[g + 1]].T @ vn[off[g]:off[g + 1]] for g in range(G)])
    np.testing.assert_allclose(S.cpu().numpy(), Sref, rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize("R", [4, 8, 20, 32])
@pytest.mark.parametrize("n_extra", [0, 1])
def test_pf2_rowpass_mma_penalty_counts(R, n_extra):
    """The tensor-core row pass is compiled per number of companion penalties (0, 1, 2): cover 0 and 1 as well
    (2 is covered by test_pf2_rowpass_both_formulations), deferred and last-iteration variants."""
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(7 + R + n_extra)
    G = 5
    sizes, off, _ = ragged(rs, G, 1, 200, R)
    N = int(off[-1])
    f = lambda *s: rs.standard_normal(size=s)  # noqa: E731
    Y, A, rho = f(N, R), rs.uniform(0.5, 1.5, size=(G, R)), rs.uniform(0.5, 2.0, size=G)
    Minv = np.stack([np.linalg.inv(m @ m.T + R * np.eye(R)) for m in f(G, R, R)])
    Wm, Delta = f(G, R, R) / np.sqrt(R), f(R, R)
    gor = np.repeat(np.arange(G), sizes)
    for deferred in (0, 1):
        for last in (0, 1):
            h = dict(pf_aux=f(N, R), pf_dual=f(N, R), l1_aux=f(N, R), l1_dual=f(N, R))
            d = {k: dev(v) for k, v in h.items()}
            pens = [(_lib.PEN_PARAFAC2, 0, 0, 0, d["pf_aux"], d["pf_dual"])]
            if n_extra:
                pens.append((_lib.PEN_L1, 0, 0.3, 0, d["l1_aux"], d["l1_dual"]))
            descs = _ops.make_descs(pens)
            x = torch.full((N, R), float("nan"), dtype=torch.float64, device="cuda")
            Wp = _ops.alloc_w(N, R, torch.float64, "cuda")
            S = torch.full((G, R, R), float("nan"), dtype=torch.float64, device="cuda")
            BtB = torch.full((G, R, R), float("nan"), dtype=torch.float64, device="cuda")
            _ops.pf2_rowpass(dev(off, torch.int64), G, R, dev(Y), dev(A), dev(rho), dev(Minv), descs, len(pens), deferred,
                             dev(Wm), dev(Delta), x if last else None, Wp if last else None, S, BtB if last else None)
            torch.cuda.synchronize()
            if deferred:
                T = np.einsum("gik,kj->gij", Wm, Delta)
                pd = np.einsum("nk,nkj->nj", h["pf_dual"], T[gor])
                dpf = h["pf_dual"] - pd
            else:
                pd, dpf = h["pf_aux"], h["pf_dual"]
            sh = pd - dpf
            if n_extra:
                sh = sh + h["l1_aux"] - h["l1_dual"]
            xr = np.einsum("nk,nkj->nj", rho[gor][:, None] * sh + Y * A[gor], Minv[gor])
            vn = xr + dpf
            np.testing.assert_allclose(d["pf_dual"].cpu().numpy(), vn, rtol=1e-11, atol=1e-10)
            if n_extra:
                v1 = xr + h["l1_dual"]
                z = np.sign(v1) * np.maximum(np.abs(v1) - 0.3 / rho[gor][:, None], 0)
                np.testing.assert_allclose(d["l1_aux"].cpu().numpy(), z, rtol=1e-11, atol=1e-10)
                np.testing.assert_allclose(d["l1_dual"].cpu().numpy(), v1 - z, rtol=1e-11, atol=1e-10)
            Sref = np.stack([vn[off[g]:off[g + 1]].T @ vn[off[g]:off[g + 1]] for g in range(G)])
            np.testing.assert_allclose(S.cpu().numpy(), Sref, rtol=1e-10, atol=1e-8)
            if last:
                np.testing.assert_allclose(x.cpu().numpy(), xr, rtol=1e-11, atol=1e-10)
                np.testing.assert_allclose(Wp[:N, :R].cpu().numpy(), xr * A[gor], rtol=1e-11, atol=1e-10)
                Bref = np.stack([xr[off[g]:off[g + 1]].T @ xr[off[g]:off[g + 1]] for g in range(G)])
                np.testing.assert_allclose(BtB.cpu().numpy(), Bref, rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize("R", [1, 3, 8, 20, 27, 32])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_pf2_gap_from_deferred_state(R, dtype):
    """b2_pf2_gap: ||V W Delta - x||^2, ||x||^2, sum|x| from the deferred PARAFAC2 state (decomposition.py:406-415)."""
    _lib, _ops, _ = _imports()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(31 + R)
    G = 9
    sizes, off, _ = ragged(rs, G, 1, 150, R)
    sizes[3] = 0
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(off[-1])
    V, x = rs.standard_normal(size=(N, R)), rs.standard_normal(size=(N, R))
    Wm, Delta = rs.standard_normal(size=(G, R, R)) / np.sqrt(R), rs.standard_normal(size=(R, R))
    d = {k: dev(v, tdt) for k, v in dict(V=V, x=x, Wm=Wm, Delta=Delta).items()}
    h = {k: v.double().cpu().numpy() for k, v in d.items()}
    out = torch.zeros(3, dtype=torch.float64, device="cuda")
    part = torch.zeros(3 * G, dtype=torch.float64, device="cuda")
    _ops.pf2_gap(d["V"], d["x"], dev(off, torch.int64), G, R, d["Wm"], d["Delta"], out, part)
    gor = np.repeat(np.arange(G), sizes)
    pd = np.einsum("nk,nkj->nj", h["V"], np.einsum("gik,kj->gij", h["Wm"], h["Delta"])[gor])
    ref = np.array([((pd - h["x"]) ** 2).sum(), (h["x"] ** 2).sum(), np.abs(h["x"]).sum()])
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-11 if dtype == "f64" else 2e-5)


@pytest.mark.parametrize("R", [2, 5, 20, 32])
def test_parafac2_polar_warm_start_equals_cold(R):
    """A Jacobi warm start from the previous call's eigenvectors must give the same W and numerator as a cold start."""
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(77 + R)
    G = 6
    V1 = [rs.standard_normal(size=(R + 10, R)) for _ in range(G)]
    V2 = [v + 1e-2 * rs.standard_normal(size=v.shape) for v in V1]
    Delta = rs.standard_normal(size=(R, R))
    rho = rs.uniform(0.5, 2.0, size=G)
    S1 = dev(np.stack([v.T @ v for v in V1]))
    S2 = dev(np.stack([v.T @ v for v in V2]))
    z = lambda: torch.zeros(G, R, R, dtype=torch.float64, device="cuda")  # noqa: E731
    Wc, nc, Ww, nw, Q = z(), z(), z(), z(), z()
    _ops.pf2_polar(S2, dev(Delta), dev(rho), G, R, Wc, nc)                       # cold reference on S2
    _ops.pf2_polar(S1, dev(Delta), dev(rho), G, R, Ww, nw, Q, warm=False)         # fills Q from S1
    _ops.pf2_polar(S2, dev(Delta), dev(rho), G, R, Ww, nw, Q, warm=True)          # warm on S2
    np.testing.assert_allclose(Ww.cpu().numpy(), Wc.cpu().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(nw.cpu().numpy(), nc.cpu().numpy(), rtol=1e-9, atol=1e-11)
    for g in range(G):  # and it is the polar factor: P = V W has orthonormal columns
        P = V2[g] @ Ww[g].cpu().numpy()
        np.testing.assert_allclose(P.T @ P, np.eye(R), atol=1e-9)


@pytest.mark.parametrize("N,K,R", [(400, 64, 9), (401, 64, 10), (333, 40, 18), (2500, 100, 20), (500, 96, 26), (450, 72, 28)])
def test_xstream_hybrid_remainder_columns(N, K, R):
    """B2_OPT_XSTREAM_HYBRID (off by default): tensor-core blocks + DFMA remainder columns must give the same Y and Z."""
    _lib, _ops, _ = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(N + R)
    Xh, X = packed_x(N, K, torch.float64, rs)
    C, Wh = rs.standard_normal(size=(K, R)), rs.standard_normal(size=(N, R))
    ws = _ops.Workspace(torch.device("cuda"), K, R, torch.float64)
    lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, 1)
    try:
        Y = torch.zeros((N, R), dtype=torch.float64, device="cuda")
        _ops.xstream_y(X, N, K, dev(C), Y, ws, _lib.VARIANT_DMMA)
        W = _ops.alloc_w(N, R, torch.float64, "cuda", _lib.VARIANT_DMMA)
        W[:N, :R] = dev(Wh)
        Z = torch.zeros((K, R), dtype=torch.float64, device="cuda")
        _ops.xstream_z(X, N, K, W, Z, ws, _lib.VARIANT_DMMA)
        torch.cuda.synchronize()
    finally:
        lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, 0)
    np.testing.assert_allclose(Y.cpu().numpy(), Xh @ C, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(Z.cpu().numpy(), Xh.T @ Wh, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("seed,skip,n", [(0, 0, 1), (1, 0, 311), (2, 3, 312), (3, 623, 5000), (4, 624, 1250), (5, 1, 100001)])
def test_device_mt19937_continues_numpy_randomstate(seed, skip, n):
    """b2_mt19937_uniform: same doubles as RandomState.uniform / random_sample, and the host generator continues the
    stream afterwards (decomposition.py:31-39 draws A, C, B_i, aux, dual from ONE RandomState)."""
    _lib, _ops, _ = _imports()
    ref = np.random.RandomState(seed)
    rs = np.random.RandomState(seed)
    if skip:
        ref.randint(0, 2 ** 31, size=skip)  # odd / block-boundary positions in the 624-word state
        rs.randint(0, 2 ** 31, size=skip)
    want = ref.uniform(size=n)
    got = _ops.mt19937_uniform(rs, n, torch.device("cuda"))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(rs.uniform(size=7), ref.uniform(size=7))          # host continues the stream
    np.testing.assert_array_equal(rs.standard_normal(size=3), ref.standard_normal(size=3))


@pytest.mark.skipif(not os.environ.get("B2_TEST_EXPERIMENTAL"),
                    reason="opt-in path written after the round's GPU budget ended: validate it first (DESIGN.md §7)")
@pytest.mark.parametrize("n_chunks,n", [(2, 1 << 20), (8, 3_000_001), (16, 5_000_000)])
def test_device_mt19937_chunked_generation(monkeypatch, n_chunks, n):
    """B2_MT_CHUNKS (opt-in): the stream cut into chunks with the host jump-ahead, one generator launch per chunk on
    its own CUDA stream — same doubles and same final generator state as one sequential walk."""
    _lib, _ops, _ = _imports()
    monkeypatch.setenv("B2_MT_CHUNKS", str(n_chunks))
    ref, rs = np.random.RandomState(5), np.random.RandomState(5)
    ref.uniform(size=13), rs.uniform(size=13)
    want = ref.uniform(size=n)
    got = _ops.mt19937_uniform(rs, n, torch.device("cuda"))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(rs.uniform(size=7), ref.uniform(size=7))


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7, 8, 13, 16, 17, 20, 21, 24, 25, 31, 32])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_parafac2_polar_warp_vs_cta_and_numpy(R, dtype):
    """B2_OPT_POLAR_WARP (default): the warp-per-slice Jacobi polar kernel against the CTA-per-slice one and against
    NumPy's SVD route (penalties.py:1233-1235), cold and warm, odd and even ranks, more slices than one wave."""
    _lib, _ops, _ = _imports()
    lib = _lib.load()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(1000 + R)
    G = 700
    V = [rs.standard_normal(size=(R + 3 + (g % 5), R)) for g in range(G)]
    Delta = rs.standard_normal(size=(R, R)) + np.eye(R)
    rho = rs.uniform(0.5, 2.0, size=G)
    S = np.stack([v.T @ v for v in V])
    Sd, Dd, rd = dev(S, tdt), dev(Delta, tdt), dev(rho, tdt)
    out = {}
    try:
        for variant in (2, 1, 0):  # registers (falls back to 1 above rank 24), shared-memory warp, CTA per slice
            lib.b2_set_option(_lib.OPT_POLAR_WARP, variant)
            Wm = torch.zeros(G, R, R, dtype=tdt, device="cuda")
            num = torch.zeros(G, R, R, dtype=torch.float64, device="cuda")
            Q = torch.zeros(G, R, R, dtype=torch.float64, device="cuda")
            _ops.pf2_polar(Sd, Dd, rd, G, R, Wm, num, Q, warm=False)
            cold = (Wm.double().cpu().numpy().copy(), num.cpu().numpy().copy())
            _ops.pf2_polar(Sd, Dd, rd, G, R, Wm, num, Q, warm=True)
            out[variant] = cold + (Wm.double().cpu().numpy(), num.cpu().numpy())
    finally:
        lib.b2_set_option(_lib.OPT_POLAR_WARP, 2)
    tol = 1e-9 if dtype == "f64" else 2e-4
    for variant in (2, 1):
        for a, b in zip(out[variant], out[0]):
            np.testing.assert_allclose(a, b, rtol=tol, atol=tol)
        np.testing.assert_allclose(out[variant][2], out[variant][0], rtol=tol, atol=tol)  # warm == cold
    out[1] = out[2]
    if dtype == "f64":
        for g in range(0, G, 37):
            U, _, Vh = np.linalg.svd(V[g] @ Delta.T, full_matrices=False)
            np.testing.assert_allclose(V[g] @ out[1][0][g], U @ Vh, atol=1e-9)
            np.testing.assert_allclose(out[1][1][g], rho[g] * (U @ Vh).T @ V[g], atol=1e-8)


# ---- "next" penalties: GeneralizedL2Penalty, UnitSimplex, TotalVariationPenalty ----------------
def test_next_penalties_prox_goldens(golden_dir):
    """Protocol calls (NumPy in, NumPy out, CUDA kernels underneath) against vectors from the reference classes."""
    from matcouply_b200 import penalties as P

    g = np.load(os.path.join(golden_dir, "operators.npz"))
    M = g["prox_in"]
    gl2 = P.GeneralizedL2Penalty(g["gl2_matrix"])
    np.testing.assert_allclose(gl2.factor_matrix_update(M, 1.3, None), g["prox_gl2"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gl2.penalty(M), float(g["gl2_value"]), rtol=1e-12)
    np.testing.assert_allclose(P.UnitSimplex().factor_matrix_update(M, 1.3, None), g["prox_simplex"], atol=5e-12)
    np.testing.assert_allclose(P.UnitSimplex().factor_matrix_update(g["prox_in_long"], 0.7, None),
                               g["prox_simplex_long"], atol=5e-12)
    np.testing.assert_allclose(P.TotalVariationPenalty(0.3).factor_matrix_update(M, 1.3, None), g["prox_tv"],
                               atol=1e-13)
    tvl1 = P.TotalVariationPenalty(0.3, l1_strength=0.2)
    np.testing.assert_allclose(tvl1.factor_matrix_update(M, 1.3, None), g["prox_tv_l1"], atol=1e-13)
    np.testing.assert_allclose(tvl1.penalty(M), float(g["tv_value"]), rtol=1e-12)
    np.testing.assert_allclose(P.TotalVariationPenalty(0.5).factor_matrix_update(g["prox_in_long"], 0.7, None),
                               g["prox_tv_long"], atol=1e-12)
    with pytest.raises(ValueError):
        P.TotalVariationPenalty(0)
    with pytest.raises(ValueError):
        P.TotalVariationPenalty(1, l1_strength=-1)
    with pytest.raises(ValueError):
        P.GeneralizedL2Penalty(np.array([[1.0, 2.0], [0.0, 1.0]]))
    with pytest.raises(ValueError):
        P.GeneralizedL2Penalty(-np.eye(3))
    assert "norm_matrix" in repr(gl2) and "validate=True" in repr(gl2)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("R", [1, 5, 20, 32])
def test_next_penalties_ragged_groups_vs_oracle(R, dtype):
    """Grouped kernels (one launch over ragged slices, per-slice rho) against the oracle, incl. dual = V - aux,
    one-row and long groups, a group too long for the shared-memory staging of the simplex kernel."""
    _lib, _ops, O = _imports()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    tol = 1e-11 if dtype == "f64" else 3e-5
    rs = np.random.RandomState(31 + R)
    sizes = [1, 2, 3, 17, 64, 65, 300, 1200 if R <= 5 else 40, 9]
    if R == 1:
        sizes.append(25000)  # > 160 KB of staging: global-memory path of the simplex kernel
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    G, N = len(sizes), int(off[-1])
    V = rs.standard_normal(size=(N, R)).cumsum(axis=0) * 0.2 + rs.standard_normal(size=(N, R))
    V = V.astype(np.float32).astype(np.float64) if dtype == "f32" else V
    rho = rs.uniform(0.3, 3.0, size=G)
    rho = rho.astype(np.float32).astype(np.float64) if dtype == "f32" else rho
    offd, rhod = dev(off, torch.int64), dev(rho, tdt)
    cut = lambda a: [a[i:j] for i, j in zip(off[:-1], off[1:])]  # noqa: E731

    aux, dual = torch.zeros(N, R, dtype=tdt, device="cuda"), dev(V, tdt)
    _ops.prox_simplex(aux, dual, offd, G, max(sizes), R)
    ref = np.concatenate([O.UnitSimplexP().prox(v, 1.0, None) for v in cut(V)], 0)
    np.testing.assert_allclose(aux.double().cpu().numpy(), ref, atol=max(tol, 5e-12))
    np.testing.assert_allclose(dual.double().cpu().numpy(), V - aux.double().cpu().numpy(), atol=tol)

    for reg, l1 in ((0.4, 0.0), (0.15, 0.3)):
        aux, dual = torch.zeros(N, R, dtype=tdt, device="cuda"), dev(V, tdt)
        _ops.prox_tv(aux, dual, offd, G, R, rhod, reg, l1)
        ref = np.concatenate([O.TotalVariationP(reg, l1).prox(v, r, None) for v, r in zip(cut(V), rho)], 0)
        np.testing.assert_allclose(aux.double().cpu().numpy(), ref, atol=tol * 10)
        np.testing.assert_allclose(dual.double().cpu().numpy(), V - aux.double().cpu().numpy(), atol=tol * 10)
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        _ops.tv_norm(dev(V, tdt), offd, G, R, out)
        np.testing.assert_allclose(out.item(), sum(np.abs(np.diff(v, axis=0)).sum() for v in cut(V)), rtol=1e-10)

    # generalized L2: equal-sized groups (J rows each), random SPD-ish norm matrix with a null space
    J, Gq = 70, 5
    L = rs.standard_normal(size=(J - 3, J))
    Mn = L.T @ L / J
    Mn = 0.5 * (Mn + Mn.T)
    Vq = rs.standard_normal(size=(Gq * J, R))
    Vq = Vq.astype(np.float32).astype(np.float64) if dtype == "f32" else Vq
    pen = O.GeneralizedL2P(Mn)
    aux, dual = torch.zeros(Gq * J, R, dtype=tdt, device="cuda"), dev(Vq, tdt)
    tmp = torch.empty_like(aux)
    _ops.prox_gl2(aux, dual, Gq, J, R, dev(pen._U, tdt), dev(pen._s, tdt), dev(rho[:Gq], tdt), tmp)
    ref = np.concatenate([pen.prox(Vq[g * J:(g + 1) * J], rho[g], None) for g in range(Gq)], 0)
    np.testing.assert_allclose(aux.double().cpu().numpy(), ref, atol=tol * 20)
    np.testing.assert_allclose(dual.double().cpu().numpy(), Vq - aux.double().cpu().numpy(), atol=tol * 20)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    _ops.quadform(dev(Mn, tdt), dev(Vq, tdt), Gq, J, R, out, tmp)
    np.testing.assert_allclose(out.item(), pen.value([Vq[g * J:(g + 1) * J] for g in range(Gq)]),
                               rtol=1e-10 if dtype == "f64" else 1e-4)


def test_tv_kernel_kkt_large():
    """KKT certificate of the CUDA TV prox on long columns (size-independent property, no oracle involved)."""
    _lib, _ops, _ = _imports()
    rs = np.random.RandomState(5)
    n, R, lam_reg, rho = 20000, 8, 0.7, 1.4
    V = np.repeat(rs.standard_normal(size=(n // 50, R)), 50, axis=0) * 2 + rs.standard_normal(size=(n, R))
    aux, dual = torch.zeros(n, R, dtype=torch.float64, device="cuda"), dev(V)
    _ops.prox_tv(aux, dual, dev(np.array([0, n]), torch.int64), 1, R, dev(np.array([rho])), lam_reg, 0.0)
    x = aux.cpu().numpy()
    lam = 2 * lam_reg / rho
    u = np.cumsum(V - x, axis=0)
    assert np.abs(u[-1]).max() < 1e-8
    assert np.all(np.abs(u[:-1]) <= lam * (1 + 1e-9) + 1e-10)
    d = np.diff(x, axis=0)
    jump = np.abs(d) > 1e-12
    np.testing.assert_allclose(u[:-1][jump], -lam * np.sign(d[jump]), atol=1e-8)


@pytest.mark.parametrize("opts", [dict(n_iter=5), dict(update_basis_matrices=False), dict(update_coordinate_matrix=False),
                                  dict(n_iter=3, update_coordinate_matrix=False), dict(n_iter=0)])
def test_parafac2_prox_options_vs_oracle(opts):
    """Parafac2(n_iter, update_basis_matrices, update_coordinate_matrix) (penalties.py:1091-1105, 1229-1248; the
    reference's own tests: tests/test_penalties.py:439-519) as protocol calls against the oracle."""
    from matcouply_b200 import penalties as P
    from oracle import aoadmm_oracle as O

    rs = np.random.RandomState(11)
    R, Js = 4, [6, 9, 4, 15, 7]
    fms = [rs.standard_normal(size=(J, R)) for J in Js]
    rhos = list(rs.uniform(0.5, 2, size=len(Js)))
    delta = rs.uniform(size=(R, R)) + 0.5 * np.eye(R)
    bases0 = [np.linalg.qr(rs.standard_normal(size=(J, R)))[0] for J in Js]
    bases, new_delta = P.Parafac2(**opts).factor_matrices_update(fms, rhos, (bases0, delta))
    ob, od = O.Parafac2P(**opts).prox_list(fms, rhos, (bases0, delta))
    for b, o in zip(bases, ob):
        np.testing.assert_allclose(b, o, atol=1e-10)
    np.testing.assert_allclose(new_delta, od, atol=1e-10)


@pytest.mark.parametrize("mma", [1, 0])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("R", [3, 4, 6, 20, 32])
def test_pf2_rowpass_single_array_companions_bit_identical(R, dtype, mma):
    """`deferred` flag bits 2 / 4 of b2_pf2_rowpass: elementwise companions carried between passes as ONE array
    T = x + dual.  Four chained passes (explicit in, T, T, explicit out) must leave bit-identical x, V, aux and dual
    as four explicit passes, for both formulations of the kernel, with a column-coupled companion left untouched."""
    _lib, _ops, _ = _imports()
    lib = _lib.load()
    tdt = torch.float64 if dtype == "f64" else torch.float32
    rs = np.random.RandomState(91 + R)
    G = 6
    sizes, off, _ = ragged(rs, G, 1, 150, R)
    N = int(off[-1])
    f = lambda *s: rs.standard_normal(size=s)  # noqa: E731
    Y, A, rho = f(N, R), rs.uniform(0.5, 1.5, size=(G, R)), rs.uniform(0.5, 2.0, size=G)
    Minv = np.stack([np.linalg.inv(m @ m.T + R * np.eye(R)) for m in f(G, R, R)]) * 0.5
    Wm, Delta = f(G, R, R) / np.sqrt(R), f(R, R) / np.sqrt(R)
    init = dict(pf_aux=f(N, R), pf_dual=f(N, R), nn_aux=f(N, R), nn_dual=f(N, R), l2_aux=f(N, R), l2_dual=f(N, R))
    offd = dev(off, torch.int64)
    results = {}
    lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, mma)
    try:
        for single in (False, True):
            d = {k: dev(v, tdt) for k, v in init.items()}
            pens = [(_lib.PEN_PARAFAC2, 0, 0, 0, d["pf_aux"], d["pf_dual"]),
                    (_lib.PEN_L1, 1, 0.2, 0, d["nn_aux"], d["nn_dual"]),
                    (_lib.PEN_L2BALL, 0, 1.0, 0, d["l2_aux"], d["l2_dual"])]
            descs = _ops.make_descs(pens)
            x = torch.zeros((N, R), dtype=tdt, device="cuda")
            Wp = _ops.alloc_w(N, R, tdt, "cuda")
            S = torch.zeros((G, R, R), dtype=tdt, device="cuda")
            BtB = torch.zeros((G, R, R), dtype=tdt, device="cuda")
            n_pass = 4
            for it in range(n_pass):
                last = it == n_pass - 1
                flags = 1 if it > 0 else 0
                if single:
                    flags |= (2 if it > 0 else 0) | (0 if last else 4)
                _ops.pf2_rowpass(offd, G, R, dev(Y, tdt), dev(A, tdt), dev(rho, tdt), dev(Minv, tdt), descs, 3, flags,
                                 dev(Wm, tdt), dev(Delta, tdt), x if last else None, Wp if last else None, S,
                                 BtB if last else None)
                # the column-coupled companion is finished by its own kernel after every pass, as in the engine
                _ops.prox_l2ball(d["l2_aux"], d["l2_dual"], offd, G, R, 1.0, 0)
            torch.cuda.synchronize()
            results[single] = {k: v.cpu().numpy().copy() for k, v in d.items() if k != "pf_aux"}
            results[single]["x"] = x.cpu().numpy().copy()
            results[single]["S"] = S.cpu().numpy().copy()
    finally:
        lib.b2_set_option(_lib.OPT_PF2_ROWPASS_MMA, _lib.PF2_ROWPASS_DEFAULT)
    for k in results[False]:
        np.testing.assert_array_equal(results[True][k], results[False][k], err_msg=k)


@pytest.mark.parametrize("K,R", [(512, 16), (100, 8), (64, 5), (1024, 8), (256, 24), (130, 16), (96, 3), (40, 32)])
@pytest.mark.parametrize("n_pen", [0, 1, 2])
@pytest.mark.parametrize("n_slices,lo,hi", [(9, 0, 150), (400, 1, 40), (150, 60, 70)])
def test_xstream_fused_local_equals_two_pass(K, R, n_pen, n_slices, lo, hi):
    """The single-read fused pass (csrc/xfused.cu: Y = X C, B-update inner loop, G_i = X_i^T B_i, Z, B_i^T B_i in ONE
    pass over X) against the two-pass composition xstream_y -> admm_local -> xstream_z and NumPy."""
    _lib, _ops, O = _imports()
    assert _ops.xstream_fused_supported(K, R, torch.float64, n_pen)
    rs = np.random.RandomState(K + 7 * R + n_pen + n_slices)
    sizes = rs.randint(lo, hi + 1, size=n_slices)
    if lo == 0:
        sizes[[0, n_slices // 2]] = 0  # empty slices
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    N = int(off[-1])
    Xh, X = packed_x(N, K, torch.float64, rs)
    C = rs.uniform(size=(K, R))
    A = rs.uniform(0.5, 1.5, size=(n_slices, R))
    A[rs.uniform(size=A.shape) < 0.1] = 0.0  # clipped entries of a non-negative A
    rho = rs.uniform(0.5, 2.0, size=n_slices)
    Ms = rs.standard_normal(size=(n_slices, 2 * R, R))
    Minv = np.stack([np.linalg.inv(m.T @ m + np.eye(R)) for m in Ms])
    cases = [(_lib.PEN_NONNEG, False, 0.0, 0.0), (_lib.PEN_L1, False, 0.2, 0.0)][:n_pen]
    aux0 = [rs.standard_normal(size=(N, R)) for _ in cases]
    dual0 = [rs.standard_normal(size=(N, R)) for _ in cases]
    gor = dev(np.repeat(np.arange(n_slices), sizes).astype(np.int32), torch.int32)
    off_d = dev(off, torch.int64)
    ws = _ops.Workspace("cuda", K, R, torch.float64)

    # two-pass composition
    aux, dual = [dev(a) for a in aux0], [dev(d) for d in dual0]
    descs = _ops.make_descs([(c[0], c[1], c[2], c[3], a, d) for c, a, d in zip(cases, aux, dual)])
    Y = torch.zeros((N, R), dtype=torch.float64, device="cuda")
    _ops.xstream_y(X, N, K, dev(C), Y, ws)
    x_ref = torch.zeros((N, R), dtype=torch.float64, device="cuda")
    W = _ops.alloc_w(N, R, torch.float64, "cuda")
    BtB_ref = torch.zeros((n_slices, R, R), dtype=torch.float64, device="cuda")
    _ops.admm_local(N, R, Y, dev(A), _lib.GROUP_INDEXED, gor, dev(rho), dev(Minv), descs, n_pen, 5, x_ref, W,
                    row_off=off_d, n_groups=n_slices, BtB_out=BtB_ref)
    Z_ref = torch.zeros((K, R), dtype=torch.float64, device="cuda")
    _ops.xstream_z(X, N, K, W, Z_ref, ws)
    ref = [x_ref] + aux + dual

    # fused pass
    aux_f, dual_f = [dev(a) for a in aux0], [dev(d) for d in dual0]
    descs_f = _ops.make_descs([(c[0], c[1], c[2], c[3], a, d) for c, a, d in zip(cases, aux_f, dual_f)])
    fws = _ops.FusedWorkspace(off, K, R, "cuda")
    assert sorted(int(g) for g in fws.sched_host.ravel() if g >= 0) == list(range(n_slices))
    x = torch.full((N, R), np.nan, dtype=torch.float64, device="cuda")
    Z = torch.full((K, R), np.nan, dtype=torch.float64, device="cuda")
    G = torch.full((n_slices, K, R), np.nan, dtype=torch.float64, device="cuda")
    BtB = torch.full((n_slices, R, R), np.nan, dtype=torch.float64, device="cuda")
    _ops.xstream_fused_local(X, N, K, off_d, n_slices, fws, dev(C), dev(A), dev(rho), dev(Minv), descs_f, n_pen, 5, x, Z,
                             G, BtB)
    torch.cuda.synchronize()
    for a, b in zip([x] + aux_f + dual_f, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-10, atol=1e-10)
    scale = max(1.0, float(Z_ref.abs().max()))
    np.testing.assert_allclose(Z.cpu().numpy(), Z_ref.cpu().numpy(), rtol=1e-10, atol=1e-11 * scale)
    np.testing.assert_allclose(BtB.cpu().numpy(), BtB_ref.cpu().numpy(), rtol=1e-10, atol=1e-10)
    xh = x.cpu().numpy()
    G_ref = np.stack([Xh[off[g]:off[g + 1]].T @ xh[off[g]:off[g + 1]] for g in range(n_slices)])
    np.testing.assert_allclose(G.cpu().numpy(), G_ref, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(G_ref).max()))
    # the A-update's right-hand side from G: diag(B_g^T X_g C2) for another C
    C2 = rs.uniform(size=(K, R))
    rhs = torch.full((n_slices, R), np.nan, dtype=torch.float64, device="cuda")
    _ops.slice_gdot(G, dev(C2), n_slices, K, R, rhs)
    rhs_ref = np.stack([np.einsum("jr,jr->r", xh[off[g]:off[g + 1]], Xh[off[g]:off[g + 1]] @ C2) for g in range(n_slices)])
    np.testing.assert_allclose(rhs.cpu().numpy(), rhs_ref, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(rhs_ref).max()))


def test_xstream_fused_local_limits():
    """Shapes whose G_i accumulators do not fit the registers (config 4: K = 2048, R = 32) are refused, not mangled."""
    _lib, _ops, O = _imports()
    assert not _ops.xstream_fused_supported(2048, 32, torch.float64, 1)
    assert not _ops.xstream_fused_supported(512, 16, torch.float32, 1)
    assert not _ops.xstream_fused_supported(512, 16, torch.float64, 3)
    assert _ops.xstream_fused_supported(512, 16, torch.float64, 1)


@pytest.mark.parametrize("N,K,R", [(1000, 512, 20), (385, 100, 17), (5000, 1030, 32), (777, 64, 24), (383, 33, 8)])
def test_xstream_y_12_consumer_warps_bit_identical(N, K, R):
    """The Y kernel with 12 consumer warps / 384-row tiles (the engine's choice from R = 17 on) contracts every row in
    the same order as the 8-warp kernel: bit-identical results, ragged tails included."""
    _lib, _ops, O = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(N + R)
    Xh, X = packed_x(N, K, torch.float64, rs)
    C = dev(rs.standard_normal(size=(K, R)))
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    outs = []
    try:
        for opt in (3, 2):
            lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, opt)
            Y = torch.full((N, R), np.nan, dtype=torch.float64, device="cuda")
            _ops.xstream_y(X, N, K, C, Y, ws, _lib.VARIANT_DMMA)
            outs.append(Y.cpu().numpy())
    finally:
        lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, 0)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_allclose(outs[1], Xh @ C.cpu().numpy(), rtol=1e-11, atol=1e-11 * np.abs(outs[1]).max())


@pytest.mark.parametrize("N,K,R", [(1000, 512, 20), (385, 100, 17), (5000, 1030, 32), (777, 64, 24), (31, 33, 8)])
def test_xstream_z_16_consumer_warps(N, K, R):
    """The Z kernel with 16 consumer warps (two per 16-k box, one partial per row half; the engine's choice from R = 17
    on) against the 8-warp kernel and NumPy (the partial sums are grouped differently: equal to round-off)."""
    _lib, _ops, O = _imports()
    lib = _lib.load()
    rs = np.random.RandomState(N + R + 1)
    Xh, X = packed_x(N, K, torch.float64, rs)
    Wh = rs.standard_normal(size=(N, R))
    W = _ops.alloc_w(N, R, torch.float64, "cuda", _lib.VARIANT_DMMA)
    W[:N, :R] = dev(Wh)
    ws = _ops.Workspace("cuda", K, R, torch.float64)
    outs = []
    try:
        for opt in (3, 2):
            lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, opt)
            Z = torch.full((K, R), np.nan, dtype=torch.float64, device="cuda")
            _ops.xstream_z(X, N, K, W, Z, ws, _lib.VARIANT_DMMA)
            outs.append(Z.cpu().numpy())
    finally:
        lib.b2_set_option(_lib.OPT_XSTREAM_HYBRID, 0)
    ref = Xh.T @ Wh
    for z in outs:
        np.testing.assert_allclose(z, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())
