"""ORACLE (host contract) — a battery of calls against the reference's HOST-side surface (result container, validation,
keyword parsing, penalty constructors), written once and run against two namespaces:

* the unmodified reference (``oracle/gen_golden_host.py`` -> ``tests/golden/host_contract.json``, build container only);
* ``matcouply_b200`` (``tests/test_host.py::test_host_contract_matches_reference``), compared entry by entry.

TEST INFRASTRUCTURE ONLY.  Every case returns either a JSON-able value or, when the call raises, the string
``"<ExceptionType>: <message>"`` — exception types AND messages are part of the drop-in contract (SURVEY.md §8b:
"exact TypeError/ValueError messages are doctested", coupled_matrices.py:79-89).  The cases restate the reference's own
host tests: tests/test_coupled_matrices.py:20-375 (container, _validate_cmf, cmf_to_*), tests/test_random.py:12-48,
tests/test_decomposition.py:455-614 (keyword parsing; `_listify`, `_parse_all_penalties`), tests/test_penalties.py
(constructor validation).
"""
import numpy as np


def _summ(v):
    """JSON-able summary of a return value: arrays by shape + a checksum, containers recursively."""
    if v is None or isinstance(v, (bool, int, str)):
        return v
    if isinstance(v, float):
        return round(v, 12)
    if isinstance(v, np.generic):
        return _summ(v.item())
    if isinstance(v, np.ndarray):
        return {"shape": list(v.shape), "sum": round(float(np.sum(v)), 9), "abs": round(float(np.sum(np.abs(v))), 9)}
    if isinstance(v, (list, tuple)):
        return [_summ(u) for u in v]
    if hasattr(v, "__iter__") and hasattr(v, "rank") and hasattr(v, "shape"):  # CoupledMatrixFactorization
        w, (A, Bs, C) = v
        return {"cmf": [_summ(w), _summ(np.asarray(A)), [_summ(np.asarray(b)) for b in Bs], _summ(np.asarray(C))],
                "shape": [list(s) for s in v.shape], "rank": int(v.rank)}
    return repr(v)


def _run(fn):
    try:
        return _summ(fn())
    except Exception as exc:  # the type and the text are what is compared
        return f"{type(exc).__name__}: {exc}"


def host_cases(cm, rnd, dec, pen):
    """cm / rnd / dec / pen: the ``coupled_matrices`` / ``random`` / ``decomposition`` / ``penalties`` modules of the
    implementation under test.  Returns {case name: outcome}."""
    rs = np.random.RandomState(42)
    rank = 3
    shapes = ((5, 4), (3, 4), (6, 4), (2, 4))
    A = rs.uniform(size=(len(shapes), rank))
    C = rs.uniform(size=(4, rank))
    Bs = [rs.uniform(size=(j, rank)) for j, _ in shapes]
    w = rs.uniform(size=rank)
    out = {}

    def case(name, fn):
        assert name not in out
        out[name] = _run(fn)

    def bs_with(first):
        c = list(Bs)
        c[0] = first
        return c

    V = cm._validate_cmf
    case("validate/ok", lambda: V((w, (A, Bs, C))))
    case("validate/ok_none_weights", lambda: V((None, (A, Bs, C))))
    case("validate/weights_scalar", lambda: V((3, (A, Bs, C))))
    case("validate/A_scalar", lambda: V((w, (1, Bs, C))))
    case("validate/A_none", lambda: V((w, (None, Bs, C))))
    case("validate/B_scalar", lambda: V((w, (A, 1, C))))
    case("validate/B_none", lambda: V((w, (A, None, C))))
    case("validate/B0_scalar", lambda: V((w, (A, bs_with(1), C))))
    case("validate/B0_none", lambda: V((w, (A, bs_with(None), C))))
    case("validate/C_scalar", lambda: V((w, (A, Bs, 1))))
    case("validate/C_none", lambda: V((w, (A, Bs, None))))
    case("validate/weights_matrix", lambda: V((np.ones((rank, rank)), (A, Bs, C))))
    case("validate/weights_too_many", lambda: V((np.ones(rank + 1), (A, Bs, C))))
    third = rs.uniform(size=(4, rank, rank))
    case("validate/A_third_order", lambda: V((w, (third, Bs, C))))
    case("validate/C_third_order", lambda: V((w, (A, Bs, third))))
    case("validate/B0_third_order", lambda: V((w, (A, bs_with(third), C))))
    vec = rs.uniform(size=rank)
    case("validate/A_vector", lambda: V((w, (vec, Bs, C))))
    case("validate/C_vector", lambda: V((w, (A, Bs, vec))))
    case("validate/B0_vector", lambda: V((w, (A, bs_with(vec), C))))
    bad_A = rs.uniform(size=(len(shapes), rank + 1))
    bad_C = rs.uniform(size=(4, rank + 1))
    case("validate/A_and_C_wrong_rank", lambda: V((w, (bad_A, Bs, bad_C))))
    case("validate/A_wrong_rank", lambda: V((w, (bad_A, Bs, C))))
    case("validate/C_wrong_rank", lambda: V((w, (A, Bs, bad_C))))
    case("validate/B0_wrong_rank", lambda: V((w, (A, bs_with(rs.uniform(size=(5, rank + 1))), C))))
    case("validate/A_too_few_rows", lambda: V((w, (A[:-1], Bs, C))))
    case("validate/too_few_B", lambda: V((w, (A, Bs[:-1], C))))
    case("validate/two_factors", lambda: V((w, (A, Bs))))

    CMF = cm.CoupledMatrixFactorization
    case("container/construct", lambda: CMF((w, (A, Bs, C))))
    case("container/getitem2", lambda: CMF((w, (A, Bs, C)))[2])
    case("container/len", lambda: len(CMF((w, (A, Bs, C)))))
    case("container/bad_construct", lambda: CMF((w, (A, Bs[:2], C))))
    for validate in (True, False):
        tag = "v" if validate else "nv"
        case(f"to_matrix/{tag}/1", lambda: cm.cmf_to_matrix((w, (A, Bs, C)), 1, validate=validate))
        case(f"to_matrix/{tag}/none_w", lambda: cm.cmf_to_matrix((None, (A, Bs, C)), 2, validate=validate))
        case(f"to_matrices/{tag}", lambda: cm.cmf_to_matrices((w, (A, Bs, C)), validate=validate))
        case(f"to_slice/{tag}", lambda: cm.cmf_to_slice((w, (A, Bs, C)), 0, validate=validate))
        case(f"to_slices/{tag}", lambda: cm.cmf_to_slices((None, (A, Bs, C)), validate=validate))
        case(f"to_tensor/{tag}", lambda: cm.cmf_to_tensor((w, (A, Bs, C)), validate=validate))
        case(f"to_vec/{tag}/pad", lambda: cm.cmf_to_vec((w, (A, Bs, C)), pad=True, validate=validate))
        case(f"to_vec/{tag}/nopad", lambda: cm.cmf_to_vec((w, (A, Bs, C)), pad=False, validate=validate))
        for mode in (0, 1, 2):
            for pad in (True, False):
                case(f"to_unfolded/{tag}/{mode}/{pad}",
                     lambda: cm.cmf_to_unfolded((w, (A, Bs, C)), mode, pad=pad, validate=validate))
    case("to_matrix/bad_cmf", lambda: cm.cmf_to_matrix((w, (bad_A, Bs, C)), 0))
    case("to_matrix/out_of_range", lambda: cm.cmf_to_matrix((w, (A, Bs, C)), 7))
    case("to_unfolded/bad_mode", lambda: cm.cmf_to_unfolded((w, (A, Bs, C)), 3))
    obj = CMF((w, (A, Bs, C)))
    case("method/to_tensor", lambda: obj.to_tensor())
    case("method/to_vec", lambda: obj.to_vec(pad=False))
    case("method/to_unfolded", lambda: obj.to_unfolded(2))
    case("method/to_matrix", lambda: obj.to_matrix(3))
    case("method/to_matrices", lambda: obj.to_matrices())

    # from_CPTensor / from_Parafac2Tensor on plain tuples (coupled_matrices.py:101-172)
    B = rs.uniform(size=(6, rank))
    case("from_cp/plain", lambda: CMF.from_CPTensor((w, (A, B, C))))
    case("from_cp/none_w", lambda: CMF.from_CPTensor((None, (A, B, C))))
    case("from_cp/shapes", lambda: CMF.from_CPTensor((w, (A, B, C)), shapes=shapes))
    case("from_cp/shapes_wrong_len", lambda: CMF.from_CPTensor((w, (A, B, C)), shapes=shapes[:3]))
    case("from_cp/shapes_wrong_K", lambda: CMF.from_CPTensor((w, (A, B, C)), shapes=[(5, 3)] * 4))
    case("from_cp/shapes_too_tall", lambda: CMF.from_CPTensor((w, (A, B, C)), shapes=[(7, 4)] * 4))
    case("from_cp/fourth_order", lambda: CMF.from_CPTensor((w, (A, B, C, C))))
    Ps = [np.linalg.qr(rs.standard_normal(size=(j, 6)))[0] for j in (8, 7, 9, 6)]
    case("from_pf2/plain", lambda: CMF.from_Parafac2Tensor((w, (A, B, C), Ps)))
    case("from_pf2/none_w", lambda: CMF.from_Parafac2Tensor((None, (A, B, C), Ps)))

    # random_coupled_matrices (random.py:9-66; tests/test_random.py:12-48)
    R = rnd.random_coupled_matrices
    case("random/default", lambda: R(shapes, rank, random_state=0))
    case("random/full", lambda: R(shapes, rank, full=True, random_state=0))
    case("random/unnormalised", lambda: R(shapes, rank, random_state=1, normalise_factors=False))
    case("random/rank_larger_than_rows", lambda: R(shapes, 5, random_state=2))
    case("random/column_mismatch", lambda: R(((3, 4), (3, 5)), 2, random_state=0))

    # keyword parsing (decomposition.py:455-467, 470-614): order and repr of the parsed penalties
    case("listify/scalar", lambda: dec._listify(3, "x"))
    case("listify/dict", lambda: dec._listify({1: 2}, "x"))
    case("listify/list", lambda: dec._listify([1, 2, 3], "x"))
    case("listify/bad_len", lambda: dec._listify([1, 2], "some_name"))

    def parse(**kw):
        base = dict(non_negative=None, lower_bound=None, upper_bound=None, l2_norm_bound=None, unimodal=None,
                    parafac2=None, l1_penalty=None, tv_penalty=None, generalized_l2_penalty=None,
                    svd="truncated_svd", regs=None, dual_init="random_uniform", aux_init="random_uniform",
                    verbose=False)
        base.update(kw)
        regs = dec._parse_all_penalties(**base)
        return [[repr(r) for r in mode] for mode in regs]

    case("parse/none", lambda: parse())
    case("parse/nonneg", lambda: parse(non_negative=True))
    case("parse/nonneg_list", lambda: parse(non_negative=[True, False, True]))
    case("parse/readme", lambda: parse(non_negative=True, l1_penalty={2: 0.1}, l2_norm_bound=[1, 1, 0],
                                       parafac2=True, unimodal={1: True}))
    case("parse/bounds", lambda: parse(lower_bound={0: -1.0}, upper_bound=[None, 2.0, 3.0], non_negative={1: True}))
    case("parse/bounds_nonneg_clamps_lower", lambda: parse(lower_bound=-1.0, non_negative=True))
    case("parse/l1_zero_strength", lambda: parse(l1_penalty=0))
    case("parse/l1_negative", lambda: parse(l1_penalty=-1.0))
    case("parse/l2ball_negative", lambda: parse(l2_norm_bound=-1.0))
    case("parse/tv_absorbs_l1", lambda: parse(tv_penalty={2: 0.3}, l1_penalty={2: 0.2}))
    case("parse/gl2", lambda: parse(generalized_l2_penalty={2: np.eye(4)}))
    case("parse/inits_forwarded", lambda: parse(non_negative=True, aux_init="zeros", dual_init="random_standard_normal"))
    case("parse/regs_not_penalty", lambda: parse(regs=[[1], [], []]))
    case("parse/regs_wrong_len", lambda: parse(regs=[[], []]))
    case("parse/regs_appended", lambda: parse(non_negative=True, regs=[[pen.L1Penalty(0.5)], [], [pen.Box(0, 1)]]))
    case("parse/bad_len", lambda: parse(non_negative=[True, False]))

    # penalty constructors and reprs (penalties.py:345-366 and the per-class validation)
    case("pen/repr_nn", lambda: repr(pen.NonNegativity()))
    case("pen/repr_box", lambda: repr(pen.Box(-1, 2, aux_init="zeros")))
    case("pen/repr_l1", lambda: repr(pen.L1Penalty(0.25, non_negativity=True)))
    case("pen/repr_l2ball", lambda: repr(pen.L2Ball(2.0)))
    case("pen/repr_unimodal", lambda: repr(pen.Unimodality(non_negativity=True)))
    case("pen/repr_pf2", lambda: repr(pen.Parafac2(n_iter=3)))
    case("pen/repr_simplex", lambda: repr(pen.UnitSimplex()))
    case("pen/repr_tv", lambda: repr(pen.TotalVariationPenalty(0.5, l1_strength=0.1)))
    case("pen/l1_negative", lambda: pen.L1Penalty(-0.1))
    case("pen/l2ball_zero", lambda: pen.L2Ball(0))
    case("pen/tv_zero", lambda: pen.TotalVariationPenalty(0))
    case("pen/tv_negative", lambda: pen.TotalVariationPenalty(-1))
    case("pen/gl2_not_square", lambda: pen.GeneralizedL2Penalty(np.ones((3, 4))))
    mats = [rs.uniform(size=(j, 4)) for j, _ in shapes]
    for name, mk in (("nn", lambda **k: pen.NonNegativity(**k)), ("l1", lambda **k: pen.L1Penalty(0.1, **k))):
        case(f"pen/{name}/aux_bad_string", lambda: mk(aux_init="nope").init_aux(mats, rank, 0, random_state=np.random.RandomState(0)))
        case(f"pen/{name}/dual_bad_string", lambda: mk(dual_init="nope").init_dual(mats, rank, 2, random_state=np.random.RandomState(0)))
        case(f"pen/{name}/aux_bad_mode", lambda: mk().init_aux(mats, rank, 3, random_state=np.random.RandomState(0)))
        case(f"pen/{name}/aux_wrong_shape_A", lambda: mk(aux_init=np.ones((2, rank))).init_aux(mats, rank, 0))
        case(f"pen/{name}/aux_wrong_shape_C", lambda: mk(aux_init=np.ones((4, rank + 1))).init_aux(mats, rank, 2))
        case(f"pen/{name}/aux_not_list_B", lambda: mk(aux_init=np.ones((5, rank))).init_aux(mats, rank, 1))
        case(f"pen/{name}/aux_list_wrong_len_B", lambda: mk(aux_init=[np.ones((5, rank))]).init_aux(mats, rank, 1))
        case(f"pen/{name}/aux_list_wrong_shape_B",
             lambda: mk(aux_init=[np.ones((j + 1, rank)) for j, _ in shapes]).init_aux(mats, rank, 1))
        case(f"pen/{name}/aux_given_ok_A", lambda: mk(aux_init=np.ones((4, rank))).init_aux(mats, rank, 0))
        case(f"pen/{name}/aux_zeros_B", lambda: mk(aux_init="zeros").init_aux(mats, rank, 1, random_state=np.random.RandomState(0)))
        case(f"pen/{name}/dual_uniform_C", lambda: mk().init_dual(mats, rank, 2, random_state=np.random.RandomState(0)))
        case(f"pen/{name}/dual_normal_B",
             lambda: mk(dual_init="random_standard_normal").init_dual(mats, rank, 1, random_state=np.random.RandomState(3)))
    case("pen/pf2/aux_mode0", lambda: pen.Parafac2().init_aux(mats, rank, 0, random_state=np.random.RandomState(0)))
    case("pen/pf2/aux_mode1", lambda: pen.Parafac2().init_aux(mats, rank, 1, random_state=np.random.RandomState(0)))
    case("pen/pf2/aux_bad_string", lambda: pen.Parafac2(aux_init="nope").init_aux(mats, rank, 1, random_state=np.random.RandomState(0)))
    case("pen/pf2/aux_not_tuple", lambda: pen.Parafac2(aux_init=[1, 2]).init_aux(mats, rank, 1, random_state=np.random.RandomState(0)))
    eyes = [np.eye(j, rank) for j, _ in shapes]
    case("pen/pf2/aux_tuple_ok", lambda: pen.Parafac2(aux_init=(eyes, np.eye(rank))).init_aux(mats, rank, 1))
    case("pen/pf2/aux_tuple_bad_delta",
         lambda: pen.Parafac2(aux_init=(eyes, np.eye(rank + 1))).init_aux(mats, rank, 1))
    case("pen/pf2/aux_tuple_not_orthogonal",
         lambda: pen.Parafac2(aux_init=([np.ones((j, rank)) for j, _ in shapes], np.eye(rank))).init_aux(mats, rank, 1))
    case("pen/pf2/aux_tuple_wrong_count", lambda: pen.Parafac2(aux_init=(eyes[:2], np.eye(rank))).init_aux(mats, rank, 1))
    case("pen/pf2/aux_tuple_wrong_rows",
         lambda: pen.Parafac2(aux_init=([np.eye(j + 1, rank) for j, _ in shapes], np.eye(rank))).init_aux(mats, rank, 1))

    # keyword names of the aliases and the state initialisers (decomposition.py:78-89)
    case("to_slice/keyword", lambda: cm.cmf_to_slice((w, (A, Bs, C)), slice_idx=2))
    case("to_matrix/keyword", lambda: cm.cmf_to_matrix((w, (A, Bs, C)), matrix_idx=2))
    regs3 = [[pen.NonNegativity()], [pen.Parafac2(), pen.L1Penalty(0.1)], [pen.Box(0, 1, aux_init="zeros")]]
    case("initialize_aux", lambda: dec.initialize_aux(mats, rank, regs3, np.random.RandomState(5)))
    case("initialize_dual", lambda: dec.initialize_dual(mats, rank, regs3, random_state=np.random.RandomState(5)))

    # the host half of the penalty protocol (penalties.py:268-343, 469-485, 1256-1324)
    aux_A, dual_A = rs.uniform(size=(4, rank)), rs.uniform(size=(4, rank))
    aux_B = [rs.uniform(size=(j, rank)) for j, _ in shapes]
    dual_B = [rs.uniform(size=(j, rank)) for j, _ in shapes]
    nn, pf2 = pen.NonNegativity(), pen.Parafac2()
    case("protocol/nn/subtract_from_aux", lambda: nn.subtract_from_aux(aux_A, dual_A))
    case("protocol/nn/subtract_from_auxes", lambda: nn.subtract_from_auxes(aux_B, dual_B))
    case("protocol/nn/aux_as_matrix", lambda: nn.aux_as_matrix(aux_A))
    case("protocol/nn/auxes_as_matrices", lambda: nn.auxes_as_matrices(aux_B))
    case("protocol/nn/penalty_matrix", lambda: nn.penalty(aux_A))
    case("protocol/nn/penalty_list", lambda: nn.penalty(aux_B))
    case("protocol/box/penalty", lambda: pen.Box(0, 1).penalty(aux_A))
    case("protocol/l2ball/penalty", lambda: pen.L2Ball(1.0).penalty(aux_B))
    case("protocol/unimodal/penalty", lambda: pen.Unimodality().penalty(aux_A))
    delta = rs.uniform(size=(rank, rank))
    case("protocol/pf2/subtract_from_auxes", lambda: pf2.subtract_from_auxes((eyes, delta), dual_B))
    case("protocol/pf2/auxes_as_matrices", lambda: pf2.auxes_as_matrices((eyes, delta)))
    case("protocol/pf2/subtract_from_aux", lambda: pf2.subtract_from_aux(aux_A, dual_A))
    case("protocol/pf2/aux_as_matrix", lambda: pf2.aux_as_matrix(aux_A))
    case("protocol/pf2/penalty_list", lambda: pf2.penalty(aux_B))
    case("protocol/pf2/penalty_matrix", lambda: pf2.penalty(aux_A))
    return out


PUBLIC_CALLABLES = {
    "decomposition": ["cmf_aoadmm", "parafac2_aoadmm", "initialize_cmf", "initialize_aux", "initialize_dual",
                      "admm_update_A", "admm_update_B", "admm_update_C", "compute_feasibility_gaps",
                      "_cmf_reconstruction_error", "_listify", "_parse_all_penalties", "_check_feasibility"],
    "penalties": ["ADMMPenalty", "MatricesPenalty", "MatrixPenalty", "RowVectorPenalty", "NonNegativity", "Box",
                  "L1Penalty", "L2Ball", "Unimodality", "Parafac2", "GeneralizedL2Penalty", "TotalVariationPenalty",
                  "UnitSimplex"],
    "coupled_matrices": ["CoupledMatrixFactorization", "cmf_to_matrix", "cmf_to_matrices", "cmf_to_slice",
                         "cmf_to_slices", "cmf_to_tensor", "cmf_to_unfolded", "cmf_to_vec", "_validate_cmf"],
    "random": ["random_coupled_matrices"],
    "data": ["get_simple_simulated_data"],
}
PROTOCOL_METHODS = ["__init__", "init_aux", "init_dual", "penalty", "subtract_from_aux", "subtract_from_auxes",
                    "aux_as_matrix", "auxes_as_matrices", "factor_matrix_row_update", "factor_matrix_update",
                    "factor_matrices_update", "from_CPTensor", "from_Parafac2Tensor", "to_matrix", "to_matrices",
                    "to_tensor", "to_unfolded", "to_vec"]


def public_signatures(modules):
    """{"module.callable[.method]": [[parameter name, default repr or "<required>"], ...]} for the public surface
    (modules: dict name -> module).  Positional order, names and defaults are the drop-in contract of a Python API."""
    import inspect

    def sig(f):
        return [[n, "<required>" if p.default is inspect.Parameter.empty else repr(p.default)]
                for n, p in inspect.signature(f).parameters.items()
                if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]

    out = {}
    for mod_name, names in PUBLIC_CALLABLES.items():
        for name in names:
            obj = getattr(modules[mod_name], name)
            if inspect.isclass(obj):
                for m in PROTOCOL_METHODS:
                    if callable(getattr(obj, m, None)):
                        out[f"{mod_name}.{name}.{m}"] = sig(getattr(obj, m))
            else:
                out[f"{mod_name}.{name}"] = sig(obj)
    return out
