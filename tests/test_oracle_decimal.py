"""Decimal128 in the oracle: pinned by the reference's own decimal vectors (binary.rs:4355-5000, transcribed in
tests/golden/decimal_kat.json), cross-checked against pyarrow's independent implementation where both follow the same rule."""
import numpy as np
import pytest

from oracle import oracle as O
from harness import load_golden
from decimal_util import col_as_py, oracle_col, oracle_nodes, parse_type

KAT = load_golden("decimal_kat.json")["cases"]


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_oracle_reproduces_reference_decimal_tests(case):
    cols = [oracle_col(c) for c in case["cols"]]
    nodes = oracle_nodes(case["rpn"])
    if "error" in case:
        with pytest.raises(O.ArrowDivideByZero):
            O.eval_expr(cols, nodes)
        return
    got = O.eval_expr(cols, nodes)
    t = parse_type(case["expected"]["type"])
    if isinstance(t, tuple):
        assert isinstance(got[0], O.Dec) and (got[0].p, got[0].s) == (t[1], t[2]), case["name"]
    else:
        assert np.asarray(got[0]).dtype == np.dtype(t), case["name"]
    assert col_as_py(got) == case["expected"]["values"], f"{case['name']} ({case['ref']})"


def _rand_dec(rng, n, p, s, null_frac=0.1):
    lim = 10 ** p - 1
    import random
    r = random.Random(int(rng.integers(0, 2 ** 31)))
    vals = [r.randint(-lim, lim) if r.random() < 0.7 else r.randint(-min(lim, 1000), min(lim, 1000)) for _ in range(n)]
    valid = rng.random(n) > null_frac
    return O.Dec(vals, p, s), valid


def test_oracle_decimal_arithmetic_vs_pyarrow():
    """+ - * follow the same published rule in Arrow C++ (precision / scale of the result, exact values)"""
    import decimal
    import pyarrow as pa
    import pyarrow.compute as pc
    rng = np.random.default_rng(3)
    decimal.getcontext().prec = 80
    for (p1, s1, p2, s2) in ((15, 2, 15, 2), (10, 0, 10, 2), (12, 4, 7, 1), (5, 5, 9, 0)):
        a, av = _rand_dec(rng, 300, p1, s1)
        b, bv = _rand_dec(rng, 300, p2, s2)
        pa_a = pa.array([decimal.Decimal(int(x)).scaleb(-s1) if k else None for x, k in zip(a, av)], pa.decimal128(p1, s1))
        pa_b = pa.array([decimal.Decimal(int(x)).scaleb(-s2) if k else None for x, k in zip(b, bv)], pa.decimal128(p2, s2))
        for op, fn in ((O.OP_PLUS, pc.add), (O.OP_MINUS, pc.subtract), (O.OP_MULTIPLY, pc.multiply)):
            got = O.eval_expr([(a, av), (b, bv)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, op, None, 0, 0)])
            ref = fn(pa_a, pa_b)
            assert (got[0].p, got[0].s) == (ref.type.precision, ref.type.scale), (op, p1, s1, p2, s2)
            ref_unscaled = [None if v is None else int(v.scaleb(ref.type.scale)) for v in ref.to_pylist()]
            assert col_as_py(got) == ref_unscaled


def test_oracle_decimal_result_types_and_errors():
    # the TPC-H revenue expression: l_extendedprice * (1 - l_discount) with Decimal128(15,2) money (benchmarks/src/tpch/mod.rs:52-122):
    # Decimal128(20,0) literal 1 (Int64 -> (20,0)) minus (15,2) -> (23,2); (15,2) * (23,2) -> (38,4)
    assert O.decimal_result_type(O.OP_MINUS, 20, 0, 15, 2)[:2] == (23, 2)
    assert O.decimal_result_type(O.OP_MULTIPLY, 15, 2, 23, 2)[:2] == (38, 4)
    big = O.Dec([10 ** 37], 38, 0)
    with pytest.raises(O.ArrowArithmeticOverflow):
        O.eval_expr([(big, None), (big, None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_MULTIPLY, None, 0, 0)])
    with pytest.raises(O.ArrowCastError):     # 12345 does not fit Decimal128(4, 0)
        O.eval_expr([(np.array([12345], np.int64), None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_CAST, 0, O.decimal_dtype(4, 0), 0, 0)])
    # rescale down rounds half away from zero (convert_to_smaller_scale_decimal)
    r = O.eval_expr([(O.Dec([125, -125, 124, -124, 135], 10, 2), None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_CAST, 0, O.decimal_dtype(10, 1), 0, 0)])
    assert col_as_py(r) == [13, -13, 12, -12, 14]
    # decimal -> int truncates
    r = O.eval_expr([(O.Dec([199, -199, 5], 10, 2), None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_CAST, 0, np.int32, 0, 0)])
    assert col_as_py(r) == [1, -1, 0]
    # NULL rows never raise
    r = O.eval_expr([(big, np.array([False])), (big, None)], [(O.E_COLUMN, 0, None, 0, 0), (O.E_COLUMN, 1, None, 0, 0), (O.E_BINARY, O.OP_MULTIPLY, None, 0, 0)])
    assert col_as_py(r) == [None]


def test_oracle_decimal_sum_is_wrapping_i128():
    # sum.rs:808-817: 99_999 + 99_999 at Decimal128(15, 2) -> 199_998
    k = np.array([7, 7], np.int64)
    keys, res = O.group_by([(k, None)], [(O.A_SUM, (O.Dec([99_999, 99_999], 15, 2), None), None)])
    assert list(res[0]["dec"]) == [199_998] and (res[0]["dec"].p, res[0]["dec"].s) == (25, 2)
    rng = np.random.default_rng(5)
    n = 5000
    g = rng.integers(0, 40, n).astype(np.int64)
    d, dv = _rand_dec(rng, n, 38, 4)
    keys, res = O.group_by([(g, None)], [(O.A_SUM, (d, dv), None), (O.A_COUNT, (d, dv), None)])
    for gi, key in enumerate(keys[0][0]):
        sel = (g == key) & dv
        tot = sum(int(x) for x in d[sel]) % (1 << 128)
        tot = tot - (1 << 128) if tot >= (1 << 127) else tot
        if sel.any():
            assert int(res[0]["dec"][gi]) == tot and res[0]["valid"][gi]
        else:
            assert not res[0]["valid"][gi]
        assert res[1]["c"][gi] == sel.sum()
