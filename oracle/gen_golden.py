"""Generate golden vectors from the UNMODIFIED reference (test infrastructure; runs only in the build container).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference (`/root/reference/src/matcouply`, pure Python) is imported as-is against the NumPy stand-in for its
missing third-party dependency `tensorly` (oracle/_tl_standin).  `/root/reference` does not exist on the GPU box, so
the vectors are committed; this script is the recipe that made them.  While generating, every case is also run
through the oracle (oracle/aoadmm_oracle.py) and the two are asserted equal to round-off — that is the pin of the
oracle against the reference.

Each trajectory file holds: the input matrices (packed rows + offsets), the keyword arguments (JSON), the factor
matrices after each of the first `n_traj` outer iterations, the final factors / aux / dual variables and all
diagnostics of a `return_errors=True` run.
"""
import json
import os
import sys
from unittest.mock import patch

sys.dont_write_bytecode = True
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache_golden")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_tl_standin"))
sys.path.insert(0, os.path.join(HERE, "_condat_standin"))  # condat_tv stand-in (TV prox = the oracle's Condat restatement)
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import matcouply.decomposition as D  # noqa: E402  (the reference)
from matcouply._unimodal_regression import unimodal_regression as ref_unimodal  # noqa: E402
from matcouply.data import get_simple_simulated_data  # noqa: E402
from matcouply import penalties as P  # noqa: E402

from oracle import aoadmm_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth(seed, I, K, J_lo, J_hi, R, kind="uniform", noise=0.2):
    """Small twins of BASELINE.json configs 1-4: truth + 20 % noise."""
    rs = np.random.RandomState(seed)
    Js = [int(j) for j in rs.randint(J_lo, J_hi + 1, size=I)]
    A = rs.uniform(size=(I, R)) + 0.1
    C = rs.uniform(size=(K, R))
    if kind == "uniform":
        Bs = [rs.uniform(size=(J, R)) for J in Js]
    elif kind == "parafac2":  # B_i = P_i Delta, clipped >= 0 is not PF2-exact but noise dominates anyway
        Delta = rs.uniform(size=(R, R)) + np.eye(R)
        Bs = []
        for J in Js:
            Q, _ = np.linalg.qr(rs.standard_normal(size=(J, R)))
            Bs.append(np.abs(Q @ Delta))
    elif kind == "gauss":  # shifting unimodal bumps (data.py:77-79 style)
        Bs = []
        for i, J in enumerate(Js):
            t = np.linspace(-10, 10, J)
            cols = [np.exp(-0.5 * (t - (-6 + 12 * r / max(R - 1, 1)) - 0.3 * i) ** 2) for r in range(R)]
            Bs.append(np.stack(cols, axis=1))
    Ms = [(B * a) @ C.T for B, a in zip(Bs, A)]
    N = [rs.standard_normal(size=M.shape) for M in Ms]
    scale = np.sqrt(sum(np.sum(M ** 2) for M in Ms)) / np.sqrt(sum(np.sum(n ** 2) for n in N))
    return [M + noise * scale * n for M, n in zip(Ms, N)]


def pack(list_of_mats):
    return np.concatenate(list_of_mats, axis=0), np.cumsum([0] + [m.shape[0] for m in list_of_mats]).astype(np.int64)


def decode_kw(kw, classes):
    """JSON-safe kwargs -> call kwargs: norm matrices as arrays, `regs_spec` -> `regs` built from `classes`."""
    kw = dict(kw)
    if isinstance(kw.get("generalized_l2_penalty"), dict):
        kw["generalized_l2_penalty"] = {int(k): np.asarray(v, dtype=np.float64)
                                        for k, v in kw["generalized_l2_penalty"].items()}
    if "regs_spec" in kw:
        kw["regs"] = O.regs_from_spec(kw.pop("regs_spec"), classes)
    return kw


def run_reference(X, rank, n_traj, kw):
    traj, orig = [], D.admm_update_A
    kw = decode_kw(kw, P)

    def spy(*a, **k):
        out = orig(*a, **k)
        _, (A, Bs, C) = out[0]
        if len(traj) < n_traj:
            traj.append((A.copy(), np.concatenate(Bs, 0).copy(), C.copy()))
        return out

    with patch("matcouply.decomposition.admm_update_A", spy):
        cmf, admm, diag = D.cmf_aoadmm([x.copy() for x in X], rank, return_errors=True, return_admm_vars=True, **kw)
    return cmf, admm, diag, traj


def flatten_aux(aux_modes, regs):
    """-> dict name -> array; mode-1 lists are packed; Parafac2 tuples split into basis/coord."""
    out = {}
    for m in range(3):
        for n, v in enumerate(aux_modes[m]):
            key = f"m{m}_r{n}"
            if isinstance(v, tuple):
                out[key + "_basis"] = np.concatenate(v[0], 0)
                out[key + "_coord"] = np.asarray(v[1])
            elif isinstance(v, list):
                out[key] = np.concatenate(v, 0)
            else:
                out[key] = np.asarray(v)
    return out


def gaps_array(gaps):
    return np.array([[g for mode in it for g in mode] for it in gaps], dtype=np.float64)


def make_case(name, X, rank, kw, n_traj=50):
    cmf, admm, diag, traj = run_reference(X, rank, n_traj, kw)
    # --- pin the oracle against the reference on this case ---
    otraj = []
    okw = dict(kw)
    if isinstance(okw.get("generalized_l2_penalty"), dict):
        okw["generalized_l2_penalty"] = {int(k): np.asarray(v) for k, v in okw["generalized_l2_penalty"].items()}
    o = O.ao_admm([x.copy() for x in X], rank, trajectory=otraj, **okw)
    assert o["n_iter"] == diag.n_iter and o["message"] == diag.message, (name, o["n_iter"], diag.n_iter)
    np.testing.assert_allclose(o["regularized_loss"], diag.regularized_loss, rtol=1e-11, atol=0)
    np.testing.assert_allclose(o["A"], cmf[1][0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(o["C"], cmf[1][2], rtol=1e-9, atol=1e-12)
    for k, (a, b, c) in enumerate(traj):
        np.testing.assert_allclose(otraj[k]["A"], a, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(np.concatenate(otraj[k]["B_is"], 0), b, rtol=1e-9, atol=1e-12)
    Xp, off = pack(X)
    data = dict(X=Xp, row_offsets=off, rank=np.int64(rank), kwargs=json.dumps(kw),
                A_traj=np.stack([t[0] for t in traj]) if traj else np.zeros((0,)),
                B_traj=np.stack([t[1] for t in traj]) if traj else np.zeros((0,)),
                C_traj=np.stack([t[2] for t in traj]) if traj else np.zeros((0,)),
                A=cmf[1][0], B=np.concatenate(cmf[1][1], 0), C=cmf[1][2],
                rec_errors=np.array(diag.rec_errors), regularized_loss=np.array(diag.regularized_loss),
                feasibility_gaps=gaps_array(diag.feasibility_gaps), n_iter=np.int64(diag.n_iter),
                message=diag.message,
                satisfied_stopping_condition=str(diag.satisfied_stopping_condition),
                satisfied_feasibility_condition=str(diag.satisfied_feasibility_condition))
    for k, v in flatten_aux(admm.auxes, None).items():
        data["aux_" + k] = v
    for k, v in flatten_aux(admm.duals, None).items():
        data["dual_" + k] = v
    np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"), **data)
    print(f"{name:28s} n_iter={diag.n_iter:4d} loss={diag.regularized_loss[-1]:.6g} rows={Xp.shape[0]} "
          f"oracle==reference OK")


def operator_goldens():
    rs = np.random.RandomState(7)
    out = {}
    # unimodal regression incl. peak index, sizes 5-500 (tests/test_unimodal_regression.py:21-41 range)
    import matcouply._unimodal_regression as U
    k = 0
    for n in [2, 3, 5, 8, 17, 50, 128, 257, 500]:
        for nn in (False, True):
            for flavour in range(3):
                if flavour == 0:
                    y = rs.standard_normal(size=(n, 4))
                elif flavour == 1:  # a noisy bump: realistic for chromatography-style data
                    t = np.linspace(-3, 3, n)[:, None]
                    y = np.exp(-t ** 2 * rs.uniform(0.5, 3, size=(1, 4))) + 0.1 * rs.standard_normal(size=(n, 4))
                else:  # many exact ties -> exercises the `<=` pooling rule
                    y = np.round(rs.standard_normal(size=(n, 4)) * 2) / 2
                fit = ref_unimodal(y, non_negativity=nn)
                peaks = []
                for r in range(4):
                    _, eL = U.prefix_isotonic_regression(y[:, r].copy(), non_negativity=nn)
                    _, eR = U.prefix_isotonic_regression(y[::-1, r].copy(), non_negativity=nn)
                    peaks.append(U._get_best_unimodality_index(eL, eR)[0])
                ofit, opeaks = O.unimodal_regression(y, nn, return_peaks=True)
                assert np.array_equal(ofit, fit) and list(opeaks) == peaks, (n, nn, flavour)
                pfit, ppeaks = O.unimodal_regression(y, nn, return_peaks=True, force_python=True)
                assert np.array_equal(pfit, fit) and list(ppeaks) == peaks
                out[f"uni{k}_y"], out[f"uni{k}_fit"] = y, fit
                out[f"uni{k}_peaks"], out[f"uni{k}_nn"] = np.array(peaks, dtype=np.int32), np.array(nn)
                k += 1
    out["uni_count"] = np.int64(k)
    # elementwise / column proxes
    M = rs.standard_normal(size=(23, 5)) * 2
    out["prox_in"] = M
    out["prox_nonneg"] = P.NonNegativity().factor_matrix_update(M, 1.3, None)
    out["prox_box"] = P.Box(-0.5, 0.7).factor_matrix_update(M, 1.3, None)
    out["prox_l1"] = P.L1Penalty(0.4).factor_matrix_update(M, 1.3, None)
    out["prox_l1_nn"] = P.L1Penalty(0.4, non_negativity=True).factor_matrix_update(M, 1.3, None)
    out["prox_l2ball"] = P.L2Ball(1.5).factor_matrix_update(M, 1.3, None)
    out["prox_l2ball_nn"] = P.L2Ball(1.5, non_negativity=True).factor_matrix_update(M, 1.3, None)
    # the "next" penalties: GeneralizedL2 and UnitSimplex from the reference classes; TV from the
    # reference class running on the condat_tv stand-in (oracle/_condat_standin)
    lap = 2 * np.eye(23) - np.eye(23, k=1) - np.eye(23, k=-1)
    lap[0, 0] = lap[-1, -1] = 1
    out["gl2_matrix"] = lap
    out["prox_gl2"] = P.GeneralizedL2Penalty(lap).factor_matrix_update(M, 1.3, None)
    out["gl2_value"] = np.float64(P.GeneralizedL2Penalty(lap).penalty(M))
    out["prox_simplex"] = P.UnitSimplex().factor_matrix_update(M, 1.3, None)
    out["prox_tv"] = P.TotalVariationPenalty(0.3).factor_matrix_update(M, 1.3, None)
    out["prox_tv_l1"] = P.TotalVariationPenalty(0.3, l1_strength=0.2).factor_matrix_update(M, 1.3, None)
    out["tv_value"] = np.float64(P.TotalVariationPenalty(0.3, l1_strength=0.2).penalty(M))
    np.testing.assert_allclose(O.GeneralizedL2P(lap).prox(M, 1.3, None), out["prox_gl2"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.UnitSimplexP().prox(M, 1.3, None), out["prox_simplex"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(O.TotalVariationP(0.3, 0.2).prox(M, 1.3, None), out["prox_tv_l1"], rtol=0, atol=1e-14)
    big = rs.standard_normal(size=(400, 3)).cumsum(axis=0) * 0.3  # long random walks: many TV segments / simplex ties
    out["prox_in_long"] = big
    out["prox_simplex_long"] = P.UnitSimplex().factor_matrix_update(big, 0.7, None)
    out["prox_tv_long"] = P.TotalVariationPenalty(0.5).factor_matrix_update(big, 0.7, None)
    # PARAFAC2 prox on a ragged list
    Js = [7, 12, 5, 9]
    fms = [rs.standard_normal(size=(J, 4)) for J in Js]
    rhos = [0.7, 1.1, 2.0, 0.4]
    delta = rs.uniform(size=(4, 4))
    pf2 = P.Parafac2()
    bases, new_delta = pf2.factor_matrices_update(fms, rhos, ([np.eye(J, 4) for J in Js], delta))
    out["pf2_in"], out["pf2_off"] = pack(fms)
    out["pf2_rhos"], out["pf2_delta_in"] = np.array(rhos), delta
    out["pf2_basis"], out["pf2_delta_out"] = np.concatenate(bases, 0), new_delta
    ob, od = O.Parafac2P().prox_list(fms, rhos, (None, delta))
    np.testing.assert_allclose(np.concatenate(ob, 0), out["pf2_basis"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(od, new_delta, rtol=1e-12, atol=1e-14)
    np.savez_compressed(os.path.join(OUT, "operators.npz"), **out)
    print(f"operators: {k} unimodal cases + prox vectors; oracle==reference OK")


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]  # optional: names of the cases to (re)generate
    if only:
        global make_case
        _make = make_case

        def make_case(name, *a, **k):  # noqa: F811
            if name in only:
                _make(name, *a, **k)
    else:
        operator_goldens()

    X0, _ = get_simple_simulated_data(noise_level=0.2, random_state=1)
    readme = dict(non_negative=True, l1_penalty={2: 0.1}, l2_norm_bound=[1, 1, 0], parafac2=True,
                  unimodal={1: True}, constant_feasibility_penalty=True, random_state=0)
    make_case("c0_readme", X0, 3, readme, n_traj=50)

    nn = dict(non_negative=True, random_state=0, n_iter_max=120)
    make_case("c1_nn_cmf", synth(1, 12, 24, 16, 16, 4), 4, nn)
    make_case("c4_nn_cmf_r8", synth(4, 6, 40, 12, 12, 8), 8, dict(nn, n_iter_max=60))

    pf2l1 = dict(non_negative=True, parafac2=True, l1_penalty={2: 0.1}, random_state=0, n_iter_max=120)
    make_case("c2_nn_pf2_l1_ragged", synth(2, 10, 20, 6, 30, 5, kind="parafac2"), 5, pf2l1)

    uni = dict(non_negative=True, parafac2=True, unimodal={1: True}, l2_norm_bound=[0, 1, 1], random_state=0,
               n_iter_max=120)
    make_case("c3_unimodal_l2ball_pf2", synth(3, 8, 16, 40, 40, 3, kind="gauss"), 3, uni)

    mixed = dict(lower_bound={1: -0.2}, upper_bound={1: 0.9}, l1_penalty={0: 0.05, 2: 0.02},
                 l2_penalty=[0.1, 0.05, 0.2], feasibility_penalty_scale=2, constant_feasibility_penalty="B",
                 random_state=3, n_iter_max=80)
    make_case("box_l1signed_l2pen", synth(5, 7, 15, 5, 14, 3), 3, mixed)

    make_case("unconstrained_l2pen", synth(6, 6, 12, 8, 11, 3), 3,
              dict(l2_penalty=0.01, random_state=1, n_iter_max=60))
    make_case("unconstrained", synth(6, 6, 12, 8, 11, 3), 3, dict(random_state=1, n_iter_max=40))

    make_case("nn_normal_init_constA", synth(7, 9, 14, 6, 10, 3), 3,
              dict(non_negative=True, aux_init="random_standard_normal", dual_init="zeros",
                   constant_feasibility_penalty="A", random_state=5, n_iter_max=60))

    make_case("freeze_C", synth(8, 6, 10, 7, 9, 2), 2,
              dict(non_negative=True, update_C=False, random_state=2, n_iter_max=30))
    make_case("nn_l2ball_mode1_only", synth(9, 8, 12, 9, 20, 4), 4,
              dict(non_negative={1: True}, l2_norm_bound={1: 0.8}, random_state=4, n_iter_max=60))


    # inner-loop convergence checks (decomposition.py:92-116): up to 20 inner iterations, stop early at inner_tol
    make_case("inner_tol_nn_pf2", synth(10, 8, 14, 6, 18, 3, kind="parafac2"), 3,
              dict(non_negative=True, parafac2=True, l2_norm_bound={2: 1.5}, inner_tol=1e-3, inner_n_iter_max=20,
                   random_state=6, n_iter_max=60))


    # ---- "next" penalties ----
    def laplacian(n):
        L = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
        L[0, 0] = L[-1, -1] = 1
        return L.tolist()

    make_case("gl2_smooth_C_nn", synth(11, 7, 18, 6, 12, 3), 3,
              dict(non_negative={0: True, 1: True}, generalized_l2_penalty={2: laplacian(18)}, random_state=2,
                   n_iter_max=60))
    make_case("gl2_smooth_B_pf2", synth(12, 6, 14, 20, 20, 3, kind="gauss"), 3,
              dict(non_negative={0: True, 2: True}, parafac2=True, generalized_l2_penalty={1: laplacian(20)},
                   random_state=3, n_iter_max=60))
    make_case("simplex_C", synth(13, 8, 15, 5, 12, 3), 3,
              dict(non_negative={0: True, 1: True}, regs_spec=[[], [], [["UnitSimplex", {}]]], random_state=1,
                   n_iter_max=60))
    make_case("simplex_B_ragged", synth(14, 8, 12, 4, 25, 4), 4,
              dict(non_negative={0: True, 2: True}, regs_spec=[[], [["UnitSimplex", {}]], []], random_state=7,
                   n_iter_max=60))
    # Parafac2 options (penalties.py:1091-1105, 1229-1248): several alternations per prox call, frozen bases / coordinates
    make_case("pf2_n_iter3_nn", synth(18, 8, 14, 6, 18, 3, kind="parafac2"), 3,
              dict(non_negative={0: True, 2: True}, regs_spec=[[], [["Parafac2", {"n_iter": 3}], ["NonNegativity", {}]], []],
                   random_state=2, n_iter_max=60))
    make_case("pf2_frozen_basis", synth(19, 7, 12, 5, 15, 3, kind="parafac2"), 3,
              dict(non_negative=True, regs_spec=[[], [["Parafac2", {"update_basis_matrices": False}]], []],
                   random_state=3, n_iter_max=40))
    make_case("pf2_frozen_coordinates", synth(20, 7, 12, 5, 15, 3, kind="parafac2"), 3,
              dict(non_negative={0: True, 2: True},
                   regs_spec=[[], [["Parafac2", {"update_coordinate_matrix": False, "n_iter": 2}]], []],
                   random_state=4, n_iter_max=40))
    # SVD-based initialisations (decomposition.py:42-53)
    make_case("init_svd_unconstrained", synth(21, 6, 12, 8, 14, 3), 3, dict(init="svd", random_state=1, n_iter_max=40))
    make_case("init_threshold_svd_nn", synth(22, 7, 14, 6, 16, 3), 3,
              dict(init="threshold_svd", non_negative=True, random_state=2, n_iter_max=60))
    make_case("tv_C_l1", synth(15, 7, 30, 6, 12, 3), 3,
              dict(non_negative={0: True, 1: True}, tv_penalty={2: 0.02}, l1_penalty={2: 0.01}, random_state=4,
                   n_iter_max=60))
    make_case("tv_B_ragged_constB", synth(16, 6, 12, 10, 40, 3, kind="gauss"), 3,
              dict(non_negative={0: True, 2: True}, tv_penalty={1: 0.05}, constant_feasibility_penalty="B",
                   random_state=5, n_iter_max=60))
    make_case("tv_A_const", synth(17, 25, 10, 5, 9, 2), 2,
              dict(non_negative={1: True, 2: True}, tv_penalty={0: 0.05}, constant_feasibility_penalty=True,
                   random_state=8, n_iter_max=60))


if __name__ == "__main__":
    main()
