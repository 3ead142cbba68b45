"""ORACLE (test infrastructure) — the seeded random problem generator of the differential tests: a random but valid
combination of ALL penalty kinds, Parafac2 options, feasibility-penalty modes, l2 penalties, inner iteration counts and
aux / dual initialisations per seed.  Shared by

* ``oracle/gen_golden_random.py``: runs the UNMODIFIED reference on seeds 0..23 -> ``tests/golden/random_configs.npz``;
* ``tests/test_oracle.py::test_oracle_matches_reference_on_random_configs`` (CPU): oracle vs those reference results;
* ``tests/test_gpu_aoadmm.py::test_random_penalty_combinations_match_oracle`` (GPU): CUDA path vs the oracle.
"""
import numpy as np

N_RANDOM_CONFIGS = 24


def random_config(seed):
    """A random but valid combination of penalties / options over all three modes (JSON-like `regs_spec`)."""
    rs = np.random.RandomState(1000 + seed)
    I, K, R = int(rs.randint(4, 9)), int(rs.randint(6, 15)), int(rs.randint(2, 6))
    equal_J = rs.rand() < 0.5
    Js = [int(rs.randint(R + 2, 20))] * I if equal_J else [int(j) for j in rs.randint(R + 2, 20, size=I)]

    def lap(n):
        L = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
        L[0, 0] = L[-1, -1] = 1
        return (0.3 * L).tolist()

    rowwise = [["NonNegativity", {}], ["Box", {"min_val": -0.1, "max_val": 0.8}], ["L1Penalty", {"reg_strength": 0.05}],
               ["L1Penalty", {"reg_strength": 0.02, "non_negativity": True}]]

    def matrixwise(n_rows):
        out = [["L2Ball", {"norm_bound": 1.2}], ["L2Ball", {"norm_bound": 0.9, "non_negativity": True}],
               ["Unimodality", {"non_negativity": bool(rs.rand() < 0.5)}], ["UnitSimplex", {}],
               ["TotalVariationPenalty", {"reg_strength": 0.03, "l1_strength": float(rs.choice([0.0, 0.01]))}]]
        if n_rows is not None:
            out.append(["GeneralizedL2Penalty", {"norm_matrix": lap(n_rows)}])
        return out

    spec, needs_const_A = [[], [], []], False
    for mode, n_rows in ((0, I), (1, Js[0] if equal_J else None), (2, K)):
        pool = rowwise + matrixwise(n_rows)
        n_pen = int(rs.randint(0, 3))
        picks = [pool[i] for i in rs.choice(len(pool), size=n_pen, replace=False)]
        if mode == 1 and rs.rand() < 0.5:
            opts = [{}, {"n_iter": 2}, {"update_coordinate_matrix": False}, {"update_basis_matrices": False}]
            picks.insert(int(rs.randint(0, len(picks) + 1)) if rs.rand() < 0.3 else 0, ["Parafac2", opts[rs.randint(4)]])
        if mode == 0 and any(p[0] not in ("NonNegativity", "Box", "L1Penalty") for p in picks):
            needs_const_A = True
        spec[mode] = picks
    const = rs.choice(["True", "A"]) if needs_const_A else rs.choice(["False", "True", "A", "B"])
    const = {"True": True, "False": False}.get(str(const), str(const))
    kw = dict(regs_spec=spec, constant_feasibility_penalty=const,
              feasibility_penalty_scale=float(rs.choice([1.0, 0.5, 2.0])),
              l2_penalty=[float(v) for v in rs.choice([0.0, 0.0, 0.05], size=3)],
              inner_n_iter_max=int(rs.choice([1, 3, 5])),
              aux_init=str(rs.choice(["random_uniform", "random_standard_normal", "zeros"])),
              dual_init=str(rs.choice(["random_uniform", "zeros"])), random_state=int(seed), n_iter_max=10,
              tol=None, absolute_tol=None)
    A, C = rs.uniform(0.2, 1.2, size=(I, R)), rs.uniform(size=(K, R))
    X = [(rs.uniform(size=(J, R)) * a) @ C.T + 0.05 * rs.standard_normal(size=(J, K)) for J, a in zip(Js, A)]
    return X, R, kw


def reference_safe(kw):
    """Delta = 0 makes the reference's Procrustes step undefined: PARAFAC2 never starts from
    aux_init="zeros" in the differential tests."""
    kw = dict(kw)
    if kw["aux_init"] == "zeros" and any(p[0] == "Parafac2" for p in kw["regs_spec"][1]):
        kw["aux_init"] = "random_uniform"
    return kw
