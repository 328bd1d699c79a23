"""Decimal128 on the GPU path: expressions, FilterExec predicates, SUM — against the reference's own decimal vectors
(tests/golden/decimal_kat.json, binary.rs:4355-5000) and against the oracle on random inputs.  Bit-exact: integer work."""
import random

import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import load_golden
from decimal_util import col_as_py, gpu_col_as_py, gpu_eval, gpu_host_col, gpu_nodes, oracle_col, oracle_nodes, parse_type

pytestmark = pytest.mark.gpu
KAT = load_golden("decimal_kat.json")["cases"]


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_gpu_reproduces_reference_decimal_tests(gpu_ctx, case):
    cols = [oracle_col(c) for c in case["cols"]]
    nodes = oracle_nodes(case["rpn"])
    if "error" in case:
        with pytest.raises(D.DfgpuError, match="Divide by zero"):
            gpu_eval(D, gpu_ctx, cols, nodes)
        return
    got, t = gpu_eval(D, gpu_ctx, cols, nodes)
    want_t = parse_type(case["expected"]["type"])
    if isinstance(want_t, tuple):
        assert t == D.decimal128(want_t[1], want_t[2]), case["name"]
    else:
        assert t == D.TYPE_OF_NP[np.dtype(want_t)], case["name"]
    assert got == case["expected"]["values"], f"{case['name']} ({case['ref']})"


def rand_dec(r, n, p, s, null_frac=0.1, small=0.3):
    lim = 10 ** p - 1
    vals = [r.randint(-min(lim, 1000), min(lim, 1000)) if r.random() < small else r.randint(-lim, lim) for _ in range(n)]
    valid = np.array([r.random() > null_frac for _ in range(n)], bool)
    return O.Dec(vals, p, s), (None if valid.all() else valid)


C0, C1 = (O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0)
def BIN(op): return (O.E_BINARY, op, None, 0, 0)
def CAST(t): return (O.E_CAST, 0, t, 0, 0)


def oracle_error_kinds(cols, nodes):
    """the error kinds present in the batch, row by row (the reference reports the first failing row; the GPU ORs the per-row flags and
    reports divide-by-zero before overflow before cast — when several kinds occur in ONE batch only that precedence can be compared)"""
    kinds = set()
    n = len(cols[0][0])
    for i in range(n):
        row = [(c[0][i:i + 1], None if c[1] is None else c[1][i:i + 1]) for c in cols]
        try:
            O.eval_expr(row, nodes)
        except O.ArrowDivideByZero:
            kinds.add("div0")
        except O.ArrowArithmeticOverflow:
            kinds.add("overflow")
        except O.ArrowCastError:
            kinds.add("cast")
    return kinds


def check_same(gpu_ctx, cols, nodes):
    """GPU == oracle, including which error is raised"""
    try:
        want = O.eval_expr(cols, nodes)
    except (O.ArrowDivideByZero, O.ArrowArithmeticOverflow, O.ArrowCastError):
        kinds = oracle_error_kinds(cols, nodes)
        kind = "div0" if "div0" in kinds else ("overflow" if "overflow" in kinds else "cast")
        with pytest.raises(D.DfgpuError, match={"div0": "Divide by zero", "overflow": "overflow", "cast": "Cast error"}[kind]):
            gpu_eval(D, gpu_ctx, cols, nodes)
        return kind
    got, t = gpu_eval(D, gpu_ctx, cols, nodes)
    if isinstance(want[0], O.Dec):
        assert t == D.decimal128(want[0].p, want[0].s)
    assert got == col_as_py(want)
    return "ok"


def test_gpu_decimal_arithmetic_random_vs_oracle(gpu_ctx):
    r = random.Random(11)
    outcomes = {}
    shapes = ((15, 2, 15, 2), (10, 0, 10, 2), (12, 4, 7, 1), (5, 5, 9, 0), (38, 10, 20, 3), (30, 0, 30, 0), (18, 6, 38, 6), (23, 2, 15, 2))
    for (p1, s1, p2, s2) in shapes:
        a, b = rand_dec(r, 700, p1, s1), rand_dec(r, 700, p2, s2)
        if (p1, s1, p2, s2) == (10, 0, 10, 2):
            # small operands and one zero divisor: the only possible error in this batch is the division by zero
            av = None if a[1] is None else a[1].copy()
            if av is not None:
                av[5] = True
            a = (O.Dec([int(x) % 1000 for x in a[0]], p1, s1), av)
            b = (O.Dec([0 if i == 5 else (int(x) % 1000) + 1 for i, x in enumerate(b[0])], p2, s2), None)
        for op in (O.OP_PLUS, O.OP_MINUS, O.OP_MULTIPLY, O.OP_DIVIDE, O.OP_MODULO):
            res = check_same(gpu_ctx, [a, b], [C0, C1, BIN(op)])
            outcomes[res] = outcomes.get(res, 0) + 1
            # the same without zero divisors / with small magnitudes so that the value path (not only the error path) is exercised
            bb = (O.Dec([int(x) if int(x) != 0 else 7 for x in b[0]], p2, s2), b[1])
            aa = (O.Dec([int(x) % 10 ** min(p1, 9) for x in a[0]], p1, s1), a[1])
            bs = (O.Dec([(int(x) % 10 ** min(p2, 9)) or 3 for x in bb[0]], p2, s2), b[1])
            res = check_same(gpu_ctx, [aa, bs], [C0, C1, BIN(op)])
            outcomes[res] = outcomes.get(res, 0) + 1
        if (p1, s1) == (p2, s2):
            for op in (O.OP_EQ, O.OP_NEQ, O.OP_LT, O.OP_LTEQ, O.OP_GT, O.OP_GTEQ, O.OP_IS_DISTINCT_FROM, O.OP_IS_NOT_DISTINCT_FROM):
                bb = (O.Dec([int(x) if r.random() < 0.5 else int(y) for x, y in zip(a[0], b[0])], p2, s2), b[1])
                assert check_same(gpu_ctx, [a, bb], [C0, C1, BIN(op)]) == "ok"
    assert outcomes.get("ok", 0) >= 40 and outcomes.get("overflow", 0) >= 1 and outcomes.get("div0", 0) >= 1, outcomes


def test_gpu_decimal_casts_vs_oracle(gpu_ctx):
    r = random.Random(12)
    d = rand_dec(r, 900, 18, 4)
    for tgt in (O.decimal_dtype(20, 6), O.decimal_dtype(38, 4), O.decimal_dtype(18, 2), O.decimal_dtype(16, 0), O.decimal_dtype(10, 1)):
        check_same(gpu_ctx, [d], [C0, CAST(tgt)])
    small = rand_dec(r, 900, 9, 3)
    for tgt in (O.decimal_dtype(9, 1), O.decimal_dtype(12, 3), O.decimal_dtype(7, 0), np.int64, np.int32, np.int16, np.float64):
        check_same(gpu_ctx, [small], [C0, CAST(tgt)])
    # rounding half away from zero on the way down; truncation toward zero into integers
    edge = (O.Dec([125, -125, 124, -124, 135, -135, 5, -5, 0, 999999999, -999999999], 10, 2), None)
    assert check_same(gpu_ctx, [edge], [C0, CAST(O.decimal_dtype(10, 1))]) == "ok"
    assert check_same(gpu_ctx, [edge], [C0, CAST(np.int64)]) == "ok"
    ints = (np.array([0, 1, -1, 123456789, -987654321, 2 ** 31 - 1, -2 ** 31], np.int64), np.array([1, 1, 1, 1, 0, 1, 1], bool))
    for tgt in (O.decimal_dtype(20, 0), O.decimal_dtype(23, 3), O.decimal_dtype(38, 20), O.decimal_dtype(12, 2)):
        check_same(gpu_ctx, [ints], [C0, CAST(tgt)])
    assert check_same(gpu_ctx, [(np.array([12345], np.int64), None)], [C0, CAST(O.decimal_dtype(4, 0))]) == "cast"
    f = (np.array([123.456, -0.5, 0.125, 1e10, -7.75, 0.0], np.float64), None)
    for tgt in (O.decimal_dtype(30, 15), O.decimal_dtype(20, 3), O.decimal_dtype(12, 0)):
        check_same(gpu_ctx, [f], [C0, CAST(tgt)])
    # a NULL row never raises
    big = (O.Dec([10 ** 37, 5], 38, 0), np.array([False, True]))
    assert check_same(gpu_ctx, [big, big], [C0, C1, BIN(O.OP_MULTIPLY)]) == "ok"
    # negative
    assert check_same(gpu_ctx, [d], [C0, (O.E_NEGATIVE, 0, None, 0, 0)]) == "ok"


def test_gpu_decimal_tpch_revenue_expression(gpu_ctx):
    """l_extendedprice * (1 - l_discount) on Decimal128(15,2) money: Decimal128(38,4) (q3.slt.part projection; tpch/mod.rs:52-122)"""
    r = random.Random(13)
    n = 20000
    price = (O.Dec([r.randint(90000, 10494950) for _ in range(n)], 15, 2), None)
    disc = (O.Dec([r.randint(0, 10) for _ in range(n)], 15, 2), None)
    nodes = [C0, (O.E_LITERAL, 0, O.decimal_dtype(20, 0), 0, 1), C1, BIN(O.OP_MINUS), BIN(O.OP_MULTIPLY)]
    want = O.eval_expr([price, disc], nodes)
    assert (want[0].p, want[0].s) == (38, 4)
    got, t = gpu_eval(D, gpu_ctx, [price, disc], nodes)
    assert t == D.decimal128(38, 4) and got == col_as_py(want)
    assert got[:3] == [int(price[0][i]) * (100 - int(disc[0][i])) for i in range(3)]


def test_gpu_filter_with_decimal_predicate(gpu_ctx):
    """TPC-H Q6 shape: l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24 over Decimal128(15,2) columns; a decimal column is carried"""
    r = random.Random(14)
    n = 30000
    disc = rand_dec(r, n, 15, 2, null_frac=0.05, small=1.0)
    disc = (O.Dec([abs(int(x)) % 11 for x in disc[0]], 15, 2), disc[1])
    qty = (O.Dec([r.randint(100, 5000) for _ in range(n)], 15, 2), None)
    price = rand_dec(r, n, 15, 2, null_frac=0.1)
    ident = (np.arange(n, dtype=np.int64), None)
    lit = lambda v: (O.E_LITERAL, 0, O.decimal_dtype(15, 2), 0, v)
    nodes = [C0, lit(5), BIN(O.OP_GTEQ), C0, lit(7), BIN(O.OP_LTEQ), BIN(O.OP_AND), C1, lit(2400), BIN(O.OP_LT), BIN(O.OP_AND)]
    cols = [disc, qty, price, ident]
    pred = O.eval_expr(cols, nodes)
    want = O.filter_batch(cols, pred)
    types = [D.decimal128(15, 2)] * 3 + [D.INT64]
    f = D.FilterHandle(gpu_ctx, types, gpu_nodes(D, nodes), None, 8192, -1)
    outs = []
    for s in range(0, n, 7000):
        hc = [gpu_host_col(D, (c[0][s:s + 7000], None if c[1] is None else c[1][s:s + 7000])) for c in cols]
        f.push_host(hc)
        outs += f.drain(host=True)
    f.finish()
    outs += f.drain(host=True)
    got = [[], [], [], []]
    for b in outs:
        for c in range(4):
            got[c] += gpu_col_as_py(D, b, c)[0]
    assert len(got[3]) == len(want[3][0]) > 100
    for c in range(4):
        assert got[c] == col_as_py(want[c])
    f.close()


@pytest.mark.parametrize("two_phase", [False, True])
def test_gpu_sum_decimal(gpu_ctx, two_phase):
    """SUM(Decimal128(p,s)) -> Decimal128(min(38, p+10), s), i128 add_wrapping (sum.rs:247-249, :316); Partial -> Final carries the state type"""
    r = random.Random(15)
    n = 40000
    g = np.array([r.randint(0, 300) for _ in range(n)], np.int64)
    d = rand_dec(r, n, 15, 2, null_frac=0.1)
    w = rand_dec(r, n, 38, 4, null_frac=0.0, small=0.0)           # sums wrap around 2^127
    w = (w[0], np.array([r.random() > 0.02 for _ in range(n)], bool))
    keys, res = O.group_by([(g, None)], [(O.A_SUM, d, None), (O.A_COUNT, d, None), (O.A_SUM, w, None)])
    want = {int(k): (None if not res[0]["valid"][i] else int(res[0]["dec"][i]), int(res[1]["c"][i]), None if not res[2]["valid"][i] else int(res[2]["dec"][i]))
            for i, k in enumerate(keys[0][0])}
    types = [D.INT64, D.decimal128(15, 2), D.decimal128(38, 4)]
    aggs = [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1), (D.AGG_SUM, 2, -1)]

    def run(handle, cols_batches):
        for hc in cols_batches:
            handle.push_host(hc)
        handle.finish()
        return handle.drain(host=True)

    batches = []
    for s in range(0, n, 9000):
        e = min(n, s + 9000)
        batches.append([D.HostColumn(g[s:e]), gpu_host_col(D, (d[0][s:e], None if d[1] is None else d[1][s:e])), gpu_host_col(D, (w[0][s:e], w[1][s:e]))])
    if not two_phase:
        h = D.AggHandle(gpu_ctx, types, [0], aggs, D.AGG_SINGLE)
        outs = run(h, batches)
    else:
        states = []
        for part in (batches[:2], batches[2:]):
            hp = D.AggHandle(gpu_ctx, types, [0], aggs, D.AGG_PARTIAL)
            states += run(hp, part)
        st_types = [D.INT64, D.decimal128(25, 2), D.INT64, D.decimal128(38, 4)]
        assert [states[0].column(i).type for i in range(4)] == st_types
        h = D.AggHandle(gpu_ctx, st_types, [0], aggs, D.AGG_FINAL)
        for b in states:
            cols = []
            for i in range(4):
                v, val = b.column_numpy(i)
                cols.append(D.HostColumn(v, val, b.column(i).type))
            h.push_host(cols)
        h.finish()
        outs = h.drain(host=True)
    got = {}
    for b in outs:
        assert b.column(1).type == D.decimal128(25, 2) and b.column(3).type == D.decimal128(38, 4)
        k = gpu_col_as_py(D, b, 0)[0]
        s1, c1, s2 = gpu_col_as_py(D, b, 1)[0], gpu_col_as_py(D, b, 2)[0], gpu_col_as_py(D, b, 3)[0]
        for i in range(len(k)):
            got[k[i]] = (s1[i], c1[i], s2[i])
    assert got == want


def test_gpu_decimal_arrow_round_trip(gpu_ctx):
    """Decimal128(15,2) through the Arrow C Data Interface: 'd:15,2' in, filter, 'd:15,2' out"""
    import decimal
    import pyarrow as pa
    vals = [decimal.Decimal("1.23"), None, decimal.Decimal("-45.60"), decimal.Decimal("99999.99"), decimal.Decimal("0.05")]
    rb = pa.record_batch([pa.array(vals, pa.decimal128(15, 2)), pa.array([1, 2, 3, 4, 5], pa.int64())], names=["m", "i"])
    lit = (O.E_LITERAL, 0, O.decimal_dtype(15, 2), 0, 100)
    f = D.FilterHandle(gpu_ctx, [D.decimal128(15, 2), D.INT64], gpu_nodes(D, [C0, lit, BIN(O.OP_GT)]), None, 8192, -1)
    f.push_arrow(rb)
    f.finish()
    outs = f.drain(host=True)
    t = pa.Table.from_batches([b.to_arrow() for b in outs])
    assert t.schema.field(0).type == pa.decimal128(15, 2)
    assert t.column(0).to_pylist() == [decimal.Decimal("1.23"), decimal.Decimal("99999.99")] and t.column(1).to_pylist() == [1, 4]
    f.close()


@pytest.mark.parametrize("nulls,mode", [(False, "single"), (True, "single"), (True, "partial_final")])
def test_fused_q3_with_decimal_money(gpu_ctx, nulls, mode):
    """the Q3 plan (tpch/plans/q3.slt.part:60-76) with the real money types: l_extendedprice, l_discount Decimal128(15,2),
    sum(l_extendedprice * (Some(1),20,0 - l_discount)) -> Decimal128(38,4); fused pipelines vs the oracle's operator chain"""
    from test_gpu_pipeline import build_lookup, oracle_filter, q3_like_tables
    from test_gpu_filter import B, C, L, to_nodes
    rng = np.random.default_rng(21)
    cust, orders, line = q3_like_tables(rng, 2000, 20_000, 80_000, nulls)
    CUT = 9200
    ct, ot = [D.INT64, D.INT64], [D.INT64, D.INT64, D.DATE32, D.INT32]
    cpred, opred = B(D.OP_EQ, C(1), L(1, np.int64)), B(D.OP_LT, C(2), L(CUT, np.int32))
    l1, _ = build_lookup(gpu_ctx, cust, ct, 0, [], pred=cpred, key_range=(1, 2000))
    l2, n2 = build_lookup(gpu_ctx, orders, ot, 0, [2, 3], pred=opred, stages=[(D.STAGE_SEMI, 1, l1)], payload_types=[D.DATE32, D.INT32],
                          n_acc_words=6, membership_filter=1)
    # lineitem with Decimal128(15,2) money
    l_key, l_price, l_disc, l_ship = line
    price = (O.Dec([int(x) for x in l_price[0]], 15, 2), l_price[1])
    disc = (O.Dec([int(x) for x in l_disc[0]], 15, 2), None)
    dline = [l_key, price, disc, l_ship]
    lt = [D.INT64, D.decimal128(15, 2), D.decimal128(15, 2), D.DATE32]
    dlit = lambda v, p, s: (O.E_LITERAL, 0, O.decimal_dtype(p, s), 0, v)
    col = lambda i: (O.E_COLUMN, i, None, 0, 0)
    # l_shipdate > CUT AND l_discount >= 0.02
    lpred = [col(3), (O.E_LITERAL, 0, np.int32, 0, CUT), BIN(O.OP_GT), col(2), dlit(2, 15, 2), BIN(O.OP_GTEQ), BIN(O.OP_AND)]
    rev = [col(1), dlit(1, 20, 0), col(2), BIN(O.OP_MINUS), BIN(O.OP_MULTIPLY)]
    p = D.Pipeline(gpu_ctx, lt, gpu_nodes(D, lpred), [(D.STAGE_INNER, 0, l2)])
    aggs = [(D.AGG_SUM, gpu_nodes(D, rev)), (D.AGG_COUNT_STAR, None), (D.AGG_COUNT, gpu_nodes(D, [col(1)]))]
    p.sink_aggregate([0, 4, 5], aggs, D.AGG_SINGLE if mode == "single" else D.AGG_PARTIAL)
    n = len(l_key[0])
    for s in range(0, n, 30_000):
        hc = [gpu_host_col(D, (c[0][s:s + 30_000], None if c[1] is None else c[1][s:s + 30_000]), t) for c, t in zip(dline, lt)]
        p.push_host(hc)
    p.finish()
    outs = p.drain(host=True)
    if mode == "partial_final":
        st_types = [outs[0].column(i).type for i in range(6)]
        assert st_types == [D.INT64, D.DATE32, D.INT32, D.decimal128(38, 4), D.INT64, D.INT64]
        h = D.AggHandle(gpu_ctx, st_types, [0, 1, 2], [(D.AGG_SUM, 3, -1), (D.AGG_COUNT_STAR, 4, -1), (D.AGG_COUNT, 5, -1)], D.AGG_FINAL)
        for b in outs:
            cols = []
            for i in range(6):
                v, val = b.column_numpy(i)
                cols.append(D.HostColumn(v, val, b.column(i).type))
            h.push_host(cols)
        h.finish()
        outs = h.drain(host=True)
    got = {}
    for b in outs:
        assert b.column(3).type == D.decimal128(38, 4)
        cols = [gpu_col_as_py(D, b, i)[0] for i in range(6)]
        for i in range(len(cols[0])):
            got[(cols[0][i], cols[1][i], cols[2][i])] = (cols[3][i], cols[4][i], cols[5][i])
    # ---- the unfused chain on the oracle ----
    fc, fo = oracle_filter(cust, cpred), oracle_filter(orders, opred)
    fl = O.filter_batch(dline, O.eval_expr(dline, lpred))
    so = O.hash_join(fc, fo, [0], [1], [1, 1, 1], [0, 2, 3], join_type=O.J_RIGHT_SEMI)
    assert len(so[0][0]) == n2
    # join on the key alone (the oracle's take() works on plain ndarrays), then gather the decimal columns by row id
    rid = (np.arange(len(fl[0][0]), dtype=np.int64), None)
    j = O.hash_join(so, [fl[0], rid], [0], [0], [1, 0, 0, 1], [0, 1, 2, 1])
    rows = j[3][0]
    jp = (O.Dec([int(fl[1][0][r]) for r in rows], 15, 2), None if fl[1][1] is None else np.asarray(fl[1][1])[rows])
    jd = (O.Dec([int(fl[2][0][r]) for r in rows], 15, 2), None)
    arg = O.eval_expr([j[0], jp, jd], rev)
    assert (arg[0].p, arg[0].s) == (38, 4)
    keys, res = O.group_by([j[0], j[1], j[2]], [(O.A_SUM, arg, None), (O.A_COUNT_STAR, None, None), (O.A_COUNT, jp, None)])
    want = {}
    for i in range(len(keys[0][0])):
        want[(int(keys[0][0][i]), int(keys[1][0][i]), int(keys[2][0][i]))] = (int(res[0]["dec"][i]) if res[0]["valid"][i] else None, int(res[1]["c"][i]), int(res[2]["c"][i]))
    assert len(want) > 500
    assert got == want
    p.close(); l2.close(); l1.close()


def test_gpu_group_by_decimal_key(gpu_ctx):
    """GROUP BY one Decimal128 column: the 128-bit value is the key (incl. -1 = all ones, 0, wide values, one NULL group)"""
    r = random.Random(16)
    n = 30000
    pool = [-1, 0, 1, 10 ** 30, -(10 ** 30), 12345, (1 << 100) + 7, -(1 << 100)] + [r.randint(-10 ** 37, 10 ** 37) for _ in range(40)]
    keys = [pool[r.randrange(len(pool))] for _ in range(n)]
    kvalid = np.array([r.random() > 0.03 for _ in range(n)], bool)
    v = np.array([r.randint(-1000, 1000) for _ in range(n)], np.int64)
    want = {}
    for i in range(n):
        k = keys[i] if kvalid[i] else None
        s, c = want.get(k, (0, 0))
        want[k] = (s + int(v[i]), c + 1)
    h = D.AggHandle(gpu_ctx, [D.decimal128(38, 0), D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT_STAR, -1, -1)], D.AGG_SINGLE)
    for s in range(0, n, 8000):
        e = min(n, s + 8000)
        h.push_host([gpu_host_col(D, (O.Dec(keys[s:e], 38, 0), kvalid[s:e])), D.HostColumn(v[s:e])])
    h.finish()
    got = {}
    for b in h.drain(host=True):
        assert b.column(0).type == D.decimal128(38, 0)
        k, s1, c1 = gpu_col_as_py(D, b, 0)[0], gpu_col_as_py(D, b, 1)[0], gpu_col_as_py(D, b, 2)[0]
        for i in range(len(k)):
            assert k[i] not in got
            got[k[i]] = (s1[i], c1[i])
    assert got == want and None in want and -1 in want
    h.close()
