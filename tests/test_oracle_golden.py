"""Pin the CPU restatement oracle against the reference's own known-answer fixtures (tests/golden/)
and against an independent engine (pyarrow.acero).  CPU only."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from harness import assert_cols_equal, col_from_list, load_golden

JT = {"Inner": O.J_INNER, "Left": O.J_LEFT, "Right": O.J_RIGHT, "Full": O.J_FULL, "LeftSemi": O.J_LEFT_SEMI, "RightSemi": O.J_RIGHT_SEMI,
      "LeftAnti": O.J_LEFT_ANTI, "RightAnti": O.J_RIGHT_ANTI, "LeftMark": O.J_LEFT_MARK, "RightMark": O.J_RIGHT_MARK}
KAT = load_golden("hash_join_kat.json")["cases"] + load_golden("hash_join_kat_extra.json")["cases"]
MISC = load_golden("misc_kat.json")


def kat_tables(case):
    # build_table_two_batches (exec.rs:3853-3861) feeds the same batch twice: case["left_repeat"/"right_repeat"]
    left = [col_from_list(list(v) * case.get("left_repeat", 1)) for _, v in case["left"]]
    right = [col_from_list(list(v) * case.get("right_repeat", 1)) for _, v in case["right"]]
    ln, rn = [n for n, _ in case["left"]], [n for n, _ in case["right"]]
    on_b = [ln.index(l) for l, _ in case["on"]]
    on_p = [rn.index(r) for _, r in case["on"]]
    jt = case["join_type"]
    if jt in ("LeftSemi", "LeftAnti"):
        side, idx = [0] * len(ln), list(range(len(ln)))
    elif jt in ("RightSemi", "RightAnti"):
        side, idx = [1] * len(rn), list(range(len(rn)))
    elif jt == "LeftMark":
        side, idx = [0] * len(ln) + [2], list(range(len(ln))) + [0]
    elif jt == "RightMark":
        side, idx = [1] * len(rn) + [2], list(range(len(rn))) + [0]
    else:
        side, idx = [0] * len(ln) + [1] * len(rn), list(range(len(ln))) + list(range(len(rn)))
    exp = []
    for c in range(len(case["header"])):
        vals = [r[c] for r in case["expected"]]
        if side[c] == 2:
            exp.append((np.array(vals, bool), None))
        else:
            exp.append(col_from_list(vals))
    return left, right, on_b, on_p, side, idx, exp


def kat_filter(case, gpu=False):
    """JoinFilter of a fixture: the shared `intermediate[0] > intermediate[1]` of the *_with_filter tests
    (prepare_join_filter, exec.rs:5556-5583), or the post-order program in filter["rpn"] (semi / anti filter tests)."""
    f = case.get("filter")
    if not f:
        return None
    rpn = f.get("rpn") or [["col", 0], ["col", 1], ["op", "gt"]]
    if gpu:
        from datafusion_b200 import capi as D
        ops = {"gt": D.OP_GT, "neq": D.OP_NEQ, "eq": D.OP_EQ, "lt": D.OP_LT}
        nodes = []
        for kind, v in rpn:
            if kind == "col":
                nodes.append((D.EXPR_COLUMN, v, 0, 0, 0, 0.0))
            elif kind == "lit_i32":
                nodes.append((D.EXPR_LITERAL, 0, D.INT32, 0, v, 0.0))
            else:
                nodes.append((D.EXPR_BINARY, ops[v], 0, 0, 0, 0.0))
    else:
        ops = {"gt": O.OP_GT, "neq": O.OP_NEQ, "eq": O.OP_EQ, "lt": O.OP_LT}
        nodes = []
        for kind, v in rpn:
            if kind == "col":
                nodes.append((O.E_COLUMN, v, None, 0, 0))
            elif kind == "lit_i32":
                nodes.append((O.E_LITERAL, 0, np.int32, 0, v))
            else:
                nodes.append((O.E_BINARY, ops[v], None, 0, 0))
    return f["col_side"], f["col_index"], nodes


# the reference's template: batch_size x perfect-hash on/off (exec.rs:2929-2962)
MATRIX = list(itertools.product([8192, 10, 5, 2, 1], [True, False]))


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_oracle_reproduces_reference_join_snapshots(case):
    left, right, on_b, on_p, side, idx, exp = kat_tables(case)
    for batch_size, phj in MATRIX:
        thr, dens = (819200, 0.0) if phj else (0, float("inf"))
        nl, nr = len(case["left"][0][1]), len(case["right"][0][1])
        fkw = {"filter": kat_filter(case)} if case.get("filter") else {}
        got = O.hash_join(left, right, on_b, on_p, side, idx, join_type=JT[case["join_type"]], **fkw, null_aware=bool(case.get("null_aware")),
                          null_equals_null=case["null_equality"] == "NullEqualsNull", batch_size=batch_size, phj_threshold=thr, phj_density=dens,
                          build_batch_rows=[nl] * case.get("left_repeat", 1), probe_batch_rows=[nr] * case.get("right_repeat", 1) if nr else None)
        assert_cols_equal(got, exp, ordered=not case["sorted"], what=f"{case['name']} bs={batch_size} phj={phj} ({case['ref']})")
        # assert_phj_used(&metrics, use_perfect_hash_join_as_possible): the ArrayMap rule of try_create_array_map (exec.rs:111-191)
        want = case.get("phj_expected", "config")
        if want is not None and len(on_b) == 1 and nl > 0 and "types" not in case:
            _, _, _, used = O.hash_join_indices([left[on_b[0]]], [right[on_p[0]]], null_equals_null=case["null_equality"] == "NullEqualsNull",
                                                phj_threshold=thr, phj_density=dens)
            assert bool(used) == bool(phj and want == "config"), f"{case['name']}: array map used={used}"


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_oracle_force_hash_collisions_is_output_invariant(case):
    # the reference's force_hash_collisions CI job (extended.yml:110-128): same results with every hash = 0
    left, right, on_b, on_p, side, idx, exp = kat_tables(case)
    fkw = {"filter": kat_filter(case)} if case.get("filter") else {}
    got = O.hash_join(left, right, on_b, on_p, side, idx, join_type=JT[case["join_type"]], phj_threshold=0, phj_density=float("inf"), force_collisions=True, **fkw,
                      null_equals_null=case["null_equality"] == "NullEqualsNull", null_aware=bool(case.get("null_aware")))
    assert_cols_equal(got, exp, ordered=not case["sorted"], what=case["name"])


def test_oracle_array_map_selection_rule():
    # exec.rs:172-179: ArrayMap iff range < threshold or density > min density (single integer key, rows < 2^32)
    k = np.arange(100, dtype=np.int64) * 3
    _, _, _, used = O.hash_join_indices([(k, None)], [(k, None)])                      # range 297 < 1024
    assert used
    k2 = np.arange(1000, dtype=np.int64) * 100
    _, _, _, used = O.hash_join_indices([(k2, None)], [(k2, None)])                    # range 99900, density 0.01
    assert not used
    k3 = np.arange(10000, dtype=np.int64) * 2
    _, _, _, used = O.hash_join_indices([(k3, None)], [(k3, None)])                    # density 0.5 > 0.15
    assert used
    _, _, _, used = O.hash_join_indices([(k3, None), (k3, None)], [(k3, None), (k3, None)])  # two keys: never
    assert not used


def test_oracle_lookup_order_example():
    m = MISC["join_lookup_doc_example"]
    b = np.array(m["build"], np.int64); p = np.array(m["probe"], np.int64)
    for phj in (True, False):
        bi, pi, _, _ = O.hash_join_indices([(b, None)], [(p, None)], phj_threshold=819200 if phj else 0, phj_density=0.0 if phj else float("inf"))
        assert bi.tolist() == m["expected_build_idx"] and pi.tolist() == m["expected_probe_idx"]
    # resumable MapOffset (join_hash_map.rs:389-484): tiny batch sizes must not change the pairs
    for bs in (1, 2, 3):
        bi, pi, _, _ = O.hash_join_indices([(b, None)], [(p, None)], batch_size=bs, phj_threshold=0, phj_density=float("inf"))
        assert bi.tolist() == m["expected_build_idx"] and pi.tolist() == m["expected_probe_idx"]


def test_oracle_perfect_hash_edge_cases():
    m = MISC["perfect_hash_negative"]
    l = np.array(m["left"][0][1], np.int64); r = np.array(m["right"][0][1], np.int64)
    for phj in (True, False):
        got = O.hash_join([(l, None)], [(r, None)], [0], [0], [0, 1], [0, 0], phj_threshold=819200 if phj else 0, phj_density=0.0 if phj else float("inf"))
        exp = [(np.array([x[0] for x in m["expected_sorted"]], np.int64), None), (np.array([x[1] for x in m["expected_sorted"]], np.int64), None)]
        assert_cols_equal(got, exp, ordered=False)
    m = MISC["perfect_hash_full_range"]
    l = np.array(m["left_i64"], np.int64); r = np.array(m["right_i64"], np.int64)
    bi, pi, _, used = O.hash_join_indices([(l, None)], [(r, None)], phj_threshold=819200, phj_density=0.0)
    assert not used and l[bi].tolist() == [m["expected_sorted"][0][0]]


def test_oracle_multi_batch_build_order():
    # exec.rs:2684-2705: the hash-map path inserts batches.iter().rev() with growing offsets and concatenates the
    # reversed list, so chains still come out in ORIGINAL build order; only the final unmatched-row pass
    # (get_final_indices_from_bit_map, utils.rs:1210-1245) walks the reversed concatenation.
    b = np.array([5, 5, 5, 5], np.int64); p = np.array([5], np.int64)
    for phj in (False, True):
        kw = dict(phj_threshold=819200, phj_density=0.0) if phj else dict(phj_threshold=0, phj_density=float("inf"))
        bi, _, _, used = O.hash_join_indices([(b, None)], [(p, None)], build_batch_rows=[2, 2], **kw)
        assert used == phj and bi.tolist() == [0, 1, 2, 3]
    b2 = np.array([1, 2, 3, 4], np.int64); p2 = np.array([9], np.int64)
    bi, pi, _, _ = O.hash_join_indices([(b2, None)], [(p2, None)], join_type=O.J_LEFT, build_batch_rows=[2, 2], phj_threshold=0, phj_density=float("inf"))
    assert bi.tolist() == [2, 3, 0, 1] and pi.tolist() == [-1] * 4       # hash map: reversed-batch concatenation
    bi, pi, _, _ = O.hash_join_indices([(b2, None)], [(p2, None)], join_type=O.J_LEFT, build_batch_rows=[2, 2])
    assert bi.tolist() == [0, 1, 2, 3]                                     # ArrayMap: concat_batches(schema, batches), exec.rs:184


def test_oracle_aggregate_some_data():
    m = MISC["aggregate_some_data"]
    a = np.concatenate([np.array(b["a"], np.uint32) for b in m["batches"]])
    v = np.concatenate([np.array(b["b"], np.float64) for b in m["batches"]])
    keys, res = O.group_by([(a, None)], [(O.A_AVG, (v, None), None)], batch_size=4)
    assert keys[0][0].tolist() == m["partial"]["a"]           # first-seen order 2,3,4
    assert res[0]["c"].tolist() == m["partial"]["count"] and res[0]["f"].tolist() == m["partial"]["sum"]
    # Final over two identical partial partitions
    pk = np.concatenate([keys[0][0], keys[0][0]]); pc = np.concatenate([res[0]["c"], res[0]["c"]]).astype(np.uint64); ps = np.concatenate([res[0]["f"], res[0]["f"]])
    fkeys, fres = O.group_by([(pk, None)], [(O.A_AVG, (pc, None), None, (ps, None))], merge=True)
    avg = O.agg_output_columns(O.A_AVG, fres[0], np.float64, state=False)[0][0]
    assert fkeys[0][0].tolist() == m["final_avg"]["a"] and avg.tolist() == m["final_avg"]["avg"]


def test_oracle_sum_count_null_state_and_wrapping():
    m = MISC["sum_null_state"]
    g1, v1 = np.array(m["g"], np.int64), col_from_list(m["v"], np.int64)
    g2, v2 = np.array(m["second_batch"]["g"], np.int64), col_from_list(m["second_batch"]["v"], np.int64)
    g = np.concatenate([g1, g2]); v = np.concatenate([v1[0], v2[0]])
    valid = np.concatenate([v1[1] if v1[1] is not None else np.ones(len(g1), bool), v2[1] if v2[1] is not None else np.ones(len(g2), bool)])
    keys, res = O.group_by([(g, None)], [(O.A_SUM, (v, valid), None), (O.A_COUNT, (v, valid), None)], batch_size=len(g1))
    exp = m["expected"]
    assert keys[0][0].tolist() == exp["g"]
    s = O.agg_output_columns(O.A_SUM, res[0], np.int64, False)[0]
    got_sum = [None if (s[1] is not None and not s[1][i]) else int(s[0][i]) for i in range(len(exp["g"]))]
    assert got_sum == exp["sum"] and res[1]["c"].tolist() == exp["count"]


def test_oracle_expressions_kat():
    m = MISC["binary_comparison"]
    a, b = np.array(m["a"], np.int32), np.array(m["b"], np.int32)
    r = O.eval_expr([(a, None), (b, None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_LT, None, 0, 0)])
    assert r[0].tolist() == m["expected"] and r[1] is None
    k = MISC["kleene"]
    a, b = col_from_list(k["a"], bool), col_from_list(k["b"], bool)
    for op, key in ((O.OP_AND, "and"), (O.OP_OR, "or")):
        v, val = O.eval_expr([a, b], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, op, None, 0, 0)])
        got = [None if (val is not None and not val[i]) else bool(v[i]) for i in range(len(v))]
        assert got == k[key]
    # float compare: -0.0 == +0.0, NaN == NaN under totalOrder (datum.rs:88-105)
    x = np.array([0.0, -0.0, np.nan, 1.0]); y = np.array([-0.0, 0.0, np.nan, np.nan])
    v, _ = O.eval_expr([(x, None), (y, None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_EQ, None, 0, 0)])
    assert v.tolist() == [True, True, True, False]
    v, _ = O.eval_expr([(x, None), (y, None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_LT, None, 0, 0)])
    assert v.tolist() == [False, False, False, True]   # 1.0 < NaN in totalOrder
    with pytest.raises(O.ArrowDivideByZero):
        O.eval_expr([(np.array([1, 2], np.int64), None), (np.array([1, 0], np.int64), None)],
                    [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_DIVIDE, None, 0, 0)])


EXPR_KAT = load_golden("expr_kat.json")["cases"]
_NP = {"int8": np.int8, "int16": np.int16, "int32": np.int32, "uint32": np.uint32, "int64": np.int64, "float32": np.float32, "float64": np.float64, "bool": bool}
_UNARY = {"not": O.E_NOT, "is_null": O.E_IS_NULL, "is_not_null": O.E_IS_NOT_NULL, "negative": O.E_NEGATIVE}
_OPS = {"eq": O.OP_EQ, "neq": O.OP_NEQ, "lt": O.OP_LT, "lteq": O.OP_LTEQ, "gt": O.OP_GT, "gteq": O.OP_GTEQ, "plus": O.OP_PLUS, "minus": O.OP_MINUS,
        "multiply": O.OP_MULTIPLY, "divide": O.OP_DIVIDE, "modulo": O.OP_MODULO, "and": O.OP_AND, "or": O.OP_OR, "is_distinct_from": O.OP_IS_DISTINCT_FROM,
        "is_not_distinct_from": O.OP_IS_NOT_DISTINCT_FROM, "bitand": O.OP_BITAND, "bitor": O.OP_BITOR, "bitxor": O.OP_BITXOR,
        "shift_left": O.OP_SHIFT_LEFT, "shift_right": O.OP_SHIFT_RIGHT}


def expr_kat_inputs(case):
    return [col_from_list(c["values"], _NP[c["type"]]) for c in case["cols"]]


def expr_kat_expected(case):
    e = case["expected"]
    return [None if v is None else (bool(v) if e["type"] == "bool" else v) for v in e["values"]]


def as_py(col, type_name):
    v, val = col
    return [None if (val is not None and not val[i]) else (bool(v[i]) if type_name == "bool" else v[i].item()) for i in range(len(v))]


@pytest.mark.parametrize("case", EXPR_KAT, ids=[c["name"] for c in EXPR_KAT])
def test_oracle_reproduces_reference_binary_expr_tests(case):
    cols = expr_kat_inputs(case)
    nodes = []
    for item in case["rpn"]:
        if item[0] == "col":
            nodes.append((O.E_COLUMN, item[1], None, 0, 0))
        elif item[0] == "lit":
            nodes.append((O.E_LITERAL, 0, _NP[item[1]], 0, item[2]))
        elif item[0] == "cast":
            nodes.append((O.E_CAST, 0, _NP[item[1]], 0, 0))
        elif item[0] in _UNARY:
            nodes.append((_UNARY[item[0]], 0, None, 0, 0))
        else:
            nodes.append((O.E_BINARY, _OPS[item[1]], None, 0, 0))
    if "error" in case:
        with pytest.raises(O.ArrowDivideByZero):
            O.eval_expr(cols, nodes)
        return
    got = O.eval_expr(cols, nodes)
    assert np.asarray(got[0]).dtype == np.dtype(_NP[case["expected"]["type"]]), case["name"]
    assert as_py(got, case["expected"]["type"]) == expr_kat_expected(case), f"{case['name']} ({case['ref']})"


# ---- independent cross-check: pyarrow.acero ------------------------------------------------
def test_oracle_vs_acero_join_and_groupby():
    import pyarrow as pa
    rng = np.random.default_rng(7)
    nb, npr = 3000, 20000
    bk = rng.integers(0, 2000, nb).astype(np.int64); bp = np.arange(nb, dtype=np.int64)
    pk = rng.integers(0, 2500, npr).astype(np.int64); pp = np.arange(npr, dtype=np.int64)
    bvalid = rng.random(nb) > 0.05; pvalid = rng.random(npr) > 0.05
    lt = pa.table({"k": pa.array(bk, mask=~bvalid), "pb": bp}); rt = pa.table({"k": pa.array(pk, mask=~pvalid), "pp": pp})
    for jt, acero in ((O.J_INNER, "inner"), (O.J_LEFT, "left outer"), (O.J_RIGHT, "right outer"), (O.J_FULL, "full outer"),
                      (O.J_LEFT_SEMI, "left semi"), (O.J_RIGHT_SEMI, "right semi"), (O.J_LEFT_ANTI, "left anti"), (O.J_RIGHT_ANTI, "right anti")):
        bi, pi, _, _ = O.hash_join_indices([(bk, bvalid)], [(pk, pvalid)], join_type=jt, phj_threshold=0, phj_density=float("inf"))
        ref = lt.join(rt, keys="k", join_type=acero, coalesce_keys=False)
        got_pb = sorted((-1 if i < 0 else int(bp[i])) for i in bi) if "right semi" not in acero and "right anti" not in acero else None
        got_pp = sorted((-1 if i < 0 else int(pp[i])) for i in pi) if "left semi" not in acero and "left anti" not in acero else None
        if got_pb is not None:
            assert got_pb == sorted(-1 if v is None else v for v in ref["pb"].to_pylist()), acero
        if got_pp is not None:
            assert got_pp == sorted(-1 if v is None else v for v in ref["pp"].to_pylist()), acero
    g = rng.integers(0, 500, 50000).astype(np.int64); v = rng.integers(-1000, 1000, 50000).astype(np.int64); vv = rng.random(50000) > 0.1
    keys, res = O.group_by([(g, None)], [(O.A_SUM, (v, vv), None), (O.A_COUNT, (v, vv), None), (O.A_MIN, (v, vv), None), (O.A_MAX, (v, vv), None)])
    t = pa.table({"g": g, "v": pa.array(v, mask=~vv)}).group_by("g").aggregate([("v", "sum"), ("v", "count"), ("v", "min"), ("v", "max")]).sort_by("g")
    o = np.argsort(keys[0][0])
    assert keys[0][0][o].tolist() == t["g"].to_pylist()
    assert res[0]["i"][o].tolist() == t["v_sum"].to_pylist() and res[1]["c"][o].tolist() == t["v_count"].to_pylist()
    assert res[2]["i"][o].tolist() == t["v_min"].to_pylist() and res[3]["i"][o].tolist() == t["v_max"].to_pylist()


def out_mapping(jt, nl, nr):
    if jt in ("LeftSemi", "LeftAnti"):
        return [0] * nl, list(range(nl))
    if jt in ("RightSemi", "RightAnti"):
        return [1] * nr, list(range(nr))
    if jt == "LeftMark":
        return [0] * nl + [2], list(range(nl)) + [0]
    if jt == "RightMark":
        return [1] * nr + [2], list(range(nr)) + [0]
    return [0] * nl + [1] * nr, list(range(nl)) + list(range(nr))


def expected_cols(rows, side):
    exp = []
    for c in range(len(side)):
        vals = [r[c] for r in rows]
        exp.append((np.array(vals, bool), None) if side[c] == 2 else col_from_list(vals))
    return exp


@pytest.mark.parametrize("jt", list(MISC["all_null_build_keys"]["expected_sorted"].keys()))
def test_oracle_all_null_build_keys(jt):
    m = MISC["all_null_build_keys"]
    left = [col_from_list(v) for _, v in m["left"]]; right = [col_from_list(v) for _, v in m["right"]]
    side, idx = out_mapping(jt, 2, 2)
    for phj in (True, False):
        got = O.hash_join(left, right, [1], [1], side, idx, join_type=JT[jt], phj_threshold=819200 if phj else 0, phj_density=0.0 if phj else float("inf"))
        assert_cols_equal(got, expected_cols(m["expected_sorted"][jt], side), ordered=False, what=f"{jt} ({m['ref']})")


def test_oracle_single_aggregate_planning_kat():
    m = MISC["single_aggregate_planning"]
    keys, res = O.group_by([(np.array(m["a_u32"], np.uint32), None)], [(O.A_SUM, (np.array(m["b_f64"], np.float64), None), None)], batch_size=2)
    s = O.agg_output_columns(O.A_SUM, res[0], np.float64, False)[0][0]
    o = np.argsort(keys[0][0])
    assert keys[0][0][o].tolist() == m["expected"]["a"] and s[o].tolist() == m["expected"]["sum"]


@pytest.mark.parametrize("name", ["skip_aggregation_after_first_batch", "skip_aggregation_after_threshold"])
def test_oracle_final_accepts_the_references_skip_aggregation_states(name):
    """The reference's Partial stream may stop aggregating and pass rows through as single-row states
    (skip_partial_aggregation_probe_*; aggregates/mod.rs:5431-5603).  Whatever it emitted, Final over those states must equal
    Single over the raw input — the contract that lets a GPU Partial (which always aggregates) feed a CPU Final and vice versa."""
    m = MISC[name]
    key = np.concatenate([np.array(b["key"], np.int32) for b in m["batches"]]); val = np.concatenate([np.array(b["val"], np.int32) for b in m["batches"]])
    sk, sres = O.group_by([(key, None)], [(O.A_COUNT, (val, None), None)])
    rp = m["reference_partial"]
    fk, fres = O.group_by([(np.array(rp["key"], np.int32), None)], [(O.A_COUNT, (np.array(rp["count"], np.int64), None), None)], merge=True)
    for keys, res in ((sk, sres), (fk, fres)):
        o = np.argsort(keys[0][0])
        assert keys[0][0][o].tolist() == m["final"]["key"] and res[0]["c"][o].tolist() == m["final"]["count"]


def test_oracle_cast_out_of_range_is_an_error():
    """DEFAULT_CAST_OPTIONS = { safe: false } (expressions/cast.rs:37-40): arrow-cast reports "Can't cast value ..." for values
    outside the integer target's range; NULL slots never raise; float -> int truncates toward zero."""
    cast = lambda col, dt: O.eval_expr([col], [(O.E_COLUMN, 0, None, 0, 0), (O.E_CAST, 0, dt, 0, 0)])
    for col, dt in (((np.array([1, -1], np.int32), None), np.uint32), ((np.array([1 << 40], np.int64), None), np.int32),
                    ((np.array([1 << 63], np.uint64), None), np.int64), ((np.array([np.nan]), None), np.int64), ((np.array([3e10]), None), np.int32)):
        with pytest.raises(O.ArrowCastError):
            cast(col, dt)
    v, val = cast((np.array([3.9, -3.9, 0.0]), None), np.int32)
    assert v.tolist() == [3, -3, 0] and val is None
    v, val = cast((np.array([5, 1 << 40], np.int64), np.array([True, False])), np.int32)    # the out-of-range value sits under a NULL
    assert v[0] == 5 and val.tolist() == [True, False]


def test_oracle_array_map_replays_the_references_unit_tests():
    """joins/array_map.rs:428-600, call by call: indices AND the resumption offsets (MapOffset) must be the reference's"""
    i32 = lambda xs: (np.array(xs, np.int64), None)
    # test_array_map_limit_offset_duplicate_elements
    build, probe = i32([1, 1, 2]), i32([1, 2])
    off, results = (0, None), []
    while off is not None:
        pi, bi, off = O.array_map_step(build, 1, 2, probe, 1, off)
        results.append((pi, bi, off))
    assert results == [([0], [0], (0, 2)), ([0], [1], (0, 0)), ([1], [2], None)]
    # test_array_map_with_limit_and_misses
    build, probe = i32([1, 2]), i32([10, 1, 2])
    pi, bi, off = O.array_map_step(build, 1, 2, probe, 1)
    assert (pi, bi, off) == ([1], [0], (2, None))
    assert O.array_map_step(build, 1, 2, probe, 1, off) == ([2], [1], None)
    # test_array_map_with_build_duplicates_and_misses
    assert O.array_map_step(i32([1, 1]), 1, 1, i32([10, 1, 20, 1]), 3) == ([1, 1, 3], [0, 1, 0], (3, 2))
    # test_array_map_rejects_large_out_of_range_probe_key (UInt64 keys; the NULL probe key never matches)
    probe = (np.array([3, (1 << 32) + 3, 11, 0], np.int64), np.array([True, True, True, False]))
    assert O.array_map_step(i32(list(range(11))), 0, 10, probe, 10) == ([0], [3], None)
    # test_array_map_i64_with_negative_and_positive_numbers (min = -5 as u64, wrapping range)
    assert O.array_map_step(i32([-5, 0, 5, -2, 3, 10]), -5, 10, i32([0, -5, 10, -1]), 10) == ([0, 1, 2], [1, 0, 5], None)


def test_oracle_join_hash_map_replays_the_references_unit_tests():
    """joins/join_hash_map.rs:517-575: NULL probe keys are skipped (valid_keys), chains come out newest row first"""
    assert O.join_hash_map_step([10, 20, 30], [10, 20, 30], np.array([True, False, True]), 8192) == ([0, 2], [0, 2], None)
    assert O.join_hash_map_step([10, 20, 10, 20], [10, 20], np.array([False, True]), 8192) == ([1, 1], [3, 1], None)
    # test_contain_hashes, via lookups: only the inserted hashes match
    pi, bi, _ = O.join_hash_map_step([10, 20, 30], [10, 11, 20, 21, 30, 31], None, 8192)
    assert pi == [0, 2, 4] and bi == [0, 1, 2]


def test_oracle_equal_rows_replays_the_references_unit_tests():
    """joins/utils.rs:4625-4870 (equal_rows_arr): the collision filter between hash lookup and emission.  String key columns of the
    reference test are dictionary-coded to integers here (a->0 .. d->3): only equality matters."""
    i = lambda xs: col_from_list(xs, np.int64)
    code = {"a": 0, "b": 1, "c": 2, "d": 3}
    left = [i([1, 2, 2, 3]), i([code[x] for x in "abcd"])]; right = [i([2, 2, 3, 4]), i([code[x] for x in "bdda"])]
    assert O.equal_rows([0, 1, 2, 3], [0, 0, 1, 2], left, right) == ([1, 3], [0, 2])                      # test_equal_rows_arr_filters_candidate_pairs
    assert O.equal_rows([0, 1, 2], [0, 1, 2], [], []) == ([], [])                                          # ..._empty_keys_returns_empty
    l, r = [i([1, None, 2, None])], [i([None, 1, 2, None])]
    assert O.equal_rows([0, 1, 2, 3], [1, 0, 2, 3], l, r, null_equals_null=False) == ([0, 2], [1, 2])      # ..._respects_null_equality
    assert O.equal_rows([0, 1, 2, 3], [1, 0, 2, 3], l, r, null_equals_null=True) == ([0, 1, 2, 3], [1, 0, 2, 3])
    f = lambda xs: (np.array(xs, np.float64).view(np.int64), None)                                         # ..._single_float_col_uses_general_path
    assert O.equal_rows([0, 1], [0, 1], [f([1.0, 2.0])], [f([1.0, 3.0])]) == ([0], [0])


def vectorized_group_values_case():
    """tests/golden/misc_kat.json vectorized_group_values_intern with the string columns dictionary-coded (only equality matters)"""
    m = MISC["vectorized_group_values_intern"]
    codes = {}

    def code(x):
        return None if x is None else codes.setdefault(x, len(codes) + 1)
    cols = [col_from_list([x for b in m["batches"] for x in b["col1"]], np.int64),
            col_from_list([code(x) for b in m["batches"] for x in b["col2"]], np.int32),
            col_from_list([code(x) for b in m["batches"] for x in b["col3"]], np.int32)]
    exp = sorted(zip(m["expected"]["col1"], [code(x) for x in m["expected"]["col2"]], [code(x) for x in m["expected"]["col3"]]),
                 key=lambda r: tuple((v is None, v or 0) for v in r))
    return cols, exp, [len(b["col1"]) for b in m["batches"]]


def group_rows(cols):
    rows = list(zip(*[[None if (c[1] is not None and not c[1][i]) else int(c[0][i]) for i in range(len(c[0]))] for c in cols]))
    return sorted(rows, key=lambda r: tuple((v is None, v or 0) for v in r))


def test_oracle_vectorized_group_values_intern_kat():
    cols, exp, sizes = vectorized_group_values_case()
    ones = (np.ones(len(cols[0][0]), np.int64), None)
    for bs in (8192, 14, 5, 1):   # one intern call per 14-row batch in the reference; the grouping must not depend on the batching
        keys, res = O.group_by(cols, [(O.A_COUNT, ones, None)], batch_size=bs)
        assert group_rows(keys) == exp
        assert int(res[0]["c"].sum()) == sum(sizes) and len(keys[0][0]) == 17


@pytest.mark.parametrize("name,threshold", [("skip_aggregation_after_first_batch", 2), ("skip_aggregation_after_threshold", 5)])
def test_oracle_reproduces_the_references_skip_partial_output(name, threshold):
    """aggregates/mod.rs:5431-5603: with probe_rows_threshold = 2 / 5 and ratio 0.1 the reference's Partial stream emits exactly these
    state rows in exactly this order (aggregated groups first, then one row per passed-through input row)."""
    m = MISC[name]
    kb = [[(np.array(b["key"], np.int32), None)] for b in m["batches"]]; ab = [(np.array(b["val"], np.int32), None) for b in m["batches"]]
    keys, state = O.partial_aggregate_with_skip(kb, ab, O.A_COUNT, probe_rows_threshold=threshold, probe_ratio_threshold=0.1)
    assert keys[0][0].tolist() == m["reference_partial"]["key"] and state[0].tolist() == m["reference_partial"]["count"]
    # the default thresholds (100 000 rows, 0.8) never trigger on this input: plain aggregation
    keys, state = O.partial_aggregate_with_skip(kb, ab, O.A_COUNT)
    assert sorted(zip(keys[0][0].tolist(), state[0].tolist())) == sorted(zip(m["final"]["key"], m["final"]["count"]))


@pytest.mark.parametrize("case", MISC["limited_batch_coalescer"]["cases"], ids=[c["name"] for c in MISC["limited_batch_coalescer"]["cases"]])
def test_oracle_limited_batch_coalescer_kat(case):
    assert O.coalesce_sizes(case["input_sizes"], case["target"], case["fetch"]) == case["expected"]


def test_oracle_join_with_hash_collisions_kat():
    """hash_join/exec.rs:5381-5510 (join_with_hash_collisions_64 / _u32): both build rows sit under BOTH probe hashes (a hand-built
    colliding JoinHashMap); lookup_join_hashmap's equality check must leave exactly (build 0, probe 0), (build 1, probe 1).
    Restated with force_collisions (every hash equal), which produces the same candidate pairs."""
    a = (np.array([10, 20], np.int64), None)
    for bs in (8192, 1):
        bi, pi, _, _ = O.hash_join_indices([a], [a], force_collisions=True, phj_threshold=0, phj_density=float("inf"), batch_size=bs)
        assert bi.tolist() == [0, 1] and pi.tolist() == [0, 1]
    # the raw-map view of the same situation: every probe hash finds the whole chain, newest row first
    pi, bi, nx = O.join_hash_map_step([7, 7], [7, 7], None, 8192)
    assert (pi, bi, nx) == ([0, 0, 1, 1], [1, 0, 1, 0], None)
    assert O.equal_rows(bi, pi, [a], [a]) == ([0, 1], [0, 1]) or sorted(zip(*O.equal_rows(bi, pi, [a], [a]))) == [(0, 0), (1, 1)]


def grouping_sets_case():
    m, sd = MISC["check_grouping_sets"], MISC["aggregate_some_data"]
    a = np.concatenate([np.array(b["a"], np.uint32) for b in sd["batches"]]); b = np.concatenate([np.array(x["b"], np.float64) for x in sd["batches"]])
    keys, copies = O.expand_grouping_sets([(a, None), (b, None)], m["masks"])
    ones = (np.ones(len(a) * copies, np.int8), None)                       # COUNT(lit(1i8))
    e = m["expected"]
    exp = sorted(zip(e["a"], e["b"], e["grouping_id"], e["count"]), key=lambda r: tuple((v is None, v or 0) for v in r))
    return keys, ones, exp


def grouping_rows(keys, counts):
    py = lambda c: [None if (c[1] is not None and not c[1][i]) else (float(c[0][i]) if np.asarray(c[0]).dtype.kind == "f" else int(c[0][i])) for i in range(len(c[0]))]
    return sorted(zip(py(keys[0]), py(keys[1]), py(keys[2]), [int(x) for x in counts]), key=lambda r: tuple((v is None, v or 0) for v in r))


def test_oracle_check_grouping_sets_kat():
    keys, ones, exp = grouping_sets_case()
    gk, res = O.group_by(keys, [(O.A_COUNT, ones, None)])
    assert grouping_rows(gk, res[0]["c"]) == exp
    # Partial -> Final over the partial states gives the same 12 rows (the test's second half)
    fk, fres = O.group_by(gk, [(O.A_COUNT, (res[0]["c"].astype(np.int64), None), None)], merge=True)
    assert grouping_rows(fk, fres[0]["c"]) == exp
