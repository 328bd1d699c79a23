"""Parity of the CUDA FilterExec / PhysicalExpr::evaluate path with the numpy restatement oracle."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import assert_cols_equal, col_from_list, gpu_filter, load_golden

pytestmark = pytest.mark.gpu
MISC = load_golden("misc_kat.json")
NP2T = {np.dtype(np.int32): D.INT32, np.dtype(np.int64): D.INT64, np.dtype(np.float64): D.FLOAT64, np.dtype(np.float32): D.FLOAT32,
        np.dtype(np.bool_): D.BOOL, np.dtype(np.uint32): D.UINT32, np.dtype(np.int8): D.INT8, np.dtype(np.uint64): D.UINT64}


def C(i): return ("col", i)
def L(v, dt, null=False): return ("lit", v, np.dtype(dt), null)
def B(op, l, r): return ("bin", op, l, r)
def U(kind, a): return ("un", kind, a)


def to_nodes(e, gpu):
    """expression tree -> post-order node lists for the C ABI (gpu=True) or the oracle"""
    out = []
    def walk(x):
        if x[0] == "col":
            out.append((D.EXPR_COLUMN, x[1], 0, 0, 0, 0.0) if gpu else (O.E_COLUMN, x[1], None, 0, 0))
        elif x[0] == "lit":
            _, v, dt, null = x
            if gpu:
                isf = dt.kind == "f"
                out.append((D.EXPR_LITERAL, 0, NP2T[dt], 1 if null else 0, 0 if (isf or null) else int(v), float(v) if (isf and not null) else 0.0))
            else:
                out.append((O.E_LITERAL, 0, dt, 1 if null else 0, v))
        elif x[0] == "bin":
            walk(x[2]); walk(x[3])
            out.append((D.EXPR_BINARY, x[1], 0, 0, 0, 0.0) if gpu else (O.E_BINARY, x[1], None, 0, 0))
        else:
            walk(x[2])
            out.append((x[1], 0, 0, 0, 0, 0.0) if gpu else (x[1], 0, None, 0, 0))
    walk(e)
    return out


def eval_gpu(ctx, cols, e):
    arr = (D.Column * len(cols))(*[D.HostColumn(v, val).c() for v, val in cols])
    keep = [D.HostColumn(v, val) for v, val in cols]
    arr = (D.Column * len(cols))(*[k.c() for k in keep])
    nodes = D.expr_nodes(to_nodes(e, True))
    import ctypes as CT
    out = CT.c_void_p()
    ctx.check(ctx.lib.dfgpu_expr_evaluate_host(ctx.h, arr, len(cols), len(cols[0][0]), nodes, len(nodes), CT.byref(out)))
    b = D.Batch(ctx, out.value)
    return b.column_numpy(0)


def test_gpu_expression_kats(gpu_ctx):
    m = MISC["binary_comparison"]
    a, b = (np.array(m["a"], np.int32), None), (np.array(m["b"], np.int32), None)
    v, val = eval_gpu(gpu_ctx, [a, b], B(D.OP_LT, C(0), C(1)))
    assert v.tolist() == m["expected"] and (val is None or val.all())
    k = MISC["kleene"]
    a, b = col_from_list(k["a"], bool), col_from_list(k["b"], bool)
    for op, key in ((D.OP_AND, "and"), (D.OP_OR, "or")):
        v, val = eval_gpu(gpu_ctx, [a, b], B(op, C(0), C(1)))
        got = [None if (val is not None and not val[i]) else bool(v[i]) for i in range(len(v))]
        assert got == k[key], key


from test_oracle_golden import EXPR_KAT, as_py, expr_kat_expected, expr_kat_inputs

_GOPS = {"eq": D.OP_EQ, "neq": D.OP_NEQ, "lt": D.OP_LT, "lteq": D.OP_LTEQ, "gt": D.OP_GT, "gteq": D.OP_GTEQ, "plus": D.OP_PLUS, "minus": D.OP_MINUS,
         "multiply": D.OP_MULTIPLY, "divide": D.OP_DIVIDE, "modulo": D.OP_MODULO, "and": D.OP_AND, "or": D.OP_OR, "is_distinct_from": D.OP_IS_DISTINCT_FROM,
         "is_not_distinct_from": D.OP_IS_NOT_DISTINCT_FROM, "bitand": D.OP_BITAND, "bitor": D.OP_BITOR, "bitxor": D.OP_BITXOR,
         "shift_left": D.OP_SHIFT_LEFT, "shift_right": D.OP_SHIFT_RIGHT}
_GT = {"int8": D.INT8, "int16": D.INT16, "int32": D.INT32, "uint32": D.UINT32, "int64": D.INT64, "float32": D.FLOAT32, "float64": D.FLOAT64, "bool": D.BOOL}
_GUNARY = {"not": D.EXPR_NOT, "is_null": D.EXPR_IS_NULL, "is_not_null": D.EXPR_IS_NOT_NULL, "negative": D.EXPR_NEGATIVE}


@pytest.mark.parametrize("case", EXPR_KAT, ids=[c["name"] for c in EXPR_KAT])
def test_gpu_reproduces_reference_binary_expr_tests(gpu_ctx, case):
    """the reference's own BinaryExpr known-answer tests (binary.rs test module), through dfgpu_expr_evaluate_host"""
    import ctypes as CT
    cols = expr_kat_inputs(case)
    keep = [D.HostColumn(v, val) for v, val in cols]
    arr = (D.Column * len(cols))(*[k.c() for k in keep])
    nodes = []
    for item in case["rpn"]:
        if item[0] == "col":
            nodes.append((D.EXPR_COLUMN, item[1], 0, 0, 0, 0.0))
        elif item[0] == "lit":
            nodes.append((D.EXPR_LITERAL, 0, _GT[item[1]], 0, item[2], 0.0))
        elif item[0] == "cast":
            nodes.append((D.EXPR_CAST, 0, _GT[item[1]], 0, 0, 0.0))
        elif item[0] in _GUNARY:
            nodes.append((_GUNARY[item[0]], 0, 0, 0, 0, 0.0))
        else:
            nodes.append((D.EXPR_BINARY, _GOPS[item[1]], 0, 0, 0, 0.0))
    na = D.expr_nodes(nodes)
    out = CT.c_void_p()
    rc = gpu_ctx.lib.dfgpu_expr_evaluate_host(gpu_ctx.h, arr, len(cols), len(cols[0][0]), na, len(nodes), CT.byref(out))
    if "error" in case:
        assert rc < 0 and case["error"].lower() in gpu_ctx.lib.dfgpu_last_error(gpu_ctx.h).decode().lower()
        return
    gpu_ctx.check(rc)
    b = D.Batch(gpu_ctx, out.value)
    assert b.column(0).type == _GT[case["expected"]["type"]], case["name"]
    assert as_py(b.column_numpy(0), case["expected"]["type"]) == expr_kat_expected(case), f"{case['name']} ({case['ref']})"


EXPRS = {
    "i64 > lit": (B(D.OP_GT, C(0), L(1 << 31, np.int64)), True),
    "(a > c) AND (b < 5 OR b IS NULL)": (B(D.OP_AND, B(D.OP_GT, C(0), L(1 << 30, np.int64)), B(D.OP_OR, B(D.OP_LT, C(1), L(5, np.int32)), U(D.EXPR_IS_NULL, C(1)))), True),
    "a + a*3 - 7 >= c (wrapping)": (B(D.OP_GTEQ, B(D.OP_MINUS, B(D.OP_PLUS, C(0), B(D.OP_MULTIPLY, C(0), L(3, np.int64))), L(7, np.int64)), L(12345, np.int64)), True),
    "b % 7 = 3": (B(D.OP_EQ, B(D.OP_MODULO, C(1), L(7, np.int32)), L(3, np.int32)), True),
    "f*2.5 <= f+1 (f64)": (B(D.OP_LTEQ, B(D.OP_MULTIPLY, C(2), L(2.5, np.float64)), B(D.OP_PLUS, C(2), L(1.0, np.float64))), True),
    "NOT(a = b64) IS DISTINCT": (B(D.OP_IS_DISTINCT_FROM, C(1), L(0, np.int32, True)), True),
    "a / 3 (value)": (B(D.OP_DIVIDE, C(0), L(3, np.int64)), False),
    "-b (value)": (U(D.EXPR_NEGATIVE, C(1)), False),
    "f - f*f (value)": (B(D.OP_MINUS, C(2), B(D.OP_MULTIPLY, C(2), C(2))), False),
}


@pytest.mark.parametrize("name", list(EXPRS.keys()))
@pytest.mark.parametrize("nulls", [False, True])
def test_gpu_expr_vs_oracle(gpu_ctx, name, nulls):
    e, is_pred = EXPRS[name]
    rng = np.random.default_rng(abs(hash(name)) % 2**31)
    n = 100_003
    a = rng.integers(0, 1 << 32, n).astype(np.int64); b = rng.integers(-50, 50, n).astype(np.int32)
    f = rng.standard_normal(n); f[::97] = np.nan; f[::89] = -0.0; f[::83] = 0.0
    mk = (lambda: rng.random(n) > 0.1) if nulls else (lambda: None)
    cols = [(a, mk()), (b, mk()), (f, mk())]
    ev, evalid = O.eval_expr(cols, to_nodes(e, False))
    gv, gvalid = eval_gpu(gpu_ctx, cols, e)
    ok = np.ones(n, bool) if evalid is None else evalid
    gok = np.ones(n, bool) if gvalid is None else gvalid
    assert np.array_equal(ok, gok), "validity"
    if ev.dtype.kind == "f":
        assert np.array_equal(ev[ok].view(np.int64), gv[ok].view(np.int64)), "float bits"   # same IEEE ops, same order: bit-exact
    else:
        assert np.array_equal(ev[ok], gv[ok])
    if is_pred:  # FilterExec over the same predicate: 128 batches of 8192 rows like config C1
        exp = O.filter_batch(cols, (ev, evalid))
        got, sizes = gpu_filter(gpu_ctx, cols, to_nodes(e, True), batch_rows=8192)
        assert_cols_equal(got, exp, ordered=True, what=name)
        assert all(s == 8192 for s in sizes[:-1]) and 0 < sizes[-1] <= 8192     # coalescer emits target-size batches


def test_gpu_filter_c1_selectivities_projection_fetch(gpu_ctx):
    # BASELINE config C1: x:int64 > c on 1M rows, 128 batches of 8192, selectivity 1/20/50/99 %
    rng = np.random.default_rng(1)
    n = 1 << 20
    x = rng.integers(0, 1 << 32, n).astype(np.int64); y = rng.integers(0, 100, n).astype(np.int32); yv = rng.random(n) > 0.1
    for sel in (0.01, 0.2, 0.5, 0.99):
        c = int(np.quantile(x, 1 - sel))
        e = B(D.OP_GT, C(0), L(c, np.int64))
        keep = x > c
        got, _ = gpu_filter(gpu_ctx, [(x, None), (y, yv)], to_nodes(e, True))
        assert_cols_equal(got, [(x[keep], None), (y[keep], yv[keep])], ordered=True, what=f"sel={sel}")
    got, _ = gpu_filter(gpu_ctx, [(x, None), (y, yv)], to_nodes(e, True), projection=[1], fetch=1000, device=True)
    assert_cols_equal(got, [(y[keep][:1000], yv[keep][:1000])], ordered=True, what="projection+fetch")


def test_gpu_divide_by_zero_is_an_error(gpu_ctx):
    a = (np.array([4, 5, 6], np.int64), None); z = (np.array([2, 0, 3], np.int64), None)
    with pytest.raises(D.DfgpuError) as ei:
        eval_gpu(gpu_ctx, [a, z], B(D.OP_DIVIDE, C(0), C(1)))
    assert "Divide by zero" in str(ei.value) and ei.value.code == -4
    zn = (np.array([2, 0, 3], np.int64), np.array([True, False, True]))       # NULL divisor slot is skipped, not an error
    v, val = eval_gpu(gpu_ctx, [a, zn], B(D.OP_DIVIDE, C(0), C(1)))
    assert val.tolist() == [True, False, True] and v[[0, 2]].tolist() == [2, 2]
    with pytest.raises(D.DfgpuError):
        D.FilterHandle(gpu_ctx, [D.INT64], to_nodes(B(D.OP_PLUS, C(0), L(1, np.int64)), True))   # non-boolean predicate (filter.rs:1355)


def test_gpu_cast_out_of_range_is_an_error(gpu_ctx):
    """same contract as test_oracle_cast_out_of_range_is_an_error (expressions/cast.rs:37-40 DEFAULT_CAST_OPTIONS)"""
    import ctypes as CT

    def cast(col, t):
        keep = D.HostColumn(*col)
        arr = (D.Column * 1)(keep.c())
        nodes = D.expr_nodes([(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_CAST, 0, t, 0, 0, 0.0)])
        out = CT.c_void_p()
        rc = gpu_ctx.lib.dfgpu_expr_evaluate_host(gpu_ctx.h, arr, 1, len(col[0]), nodes, 2, CT.byref(out))
        if rc < 0:
            return rc, gpu_ctx.lib.dfgpu_last_error(gpu_ctx.h).decode()
        return 0, D.Batch(gpu_ctx, out.value).column_numpy(0)

    for col, t in (((np.array([1, -1], np.int32), None), D.UINT32), ((np.array([1 << 40], np.int64), None), D.INT32),
                   ((np.array([1 << 63], np.uint64), None), D.INT64), ((np.array([np.nan]), None), D.INT64), ((np.array([3e10]), None), D.INT32)):
        rc, msg = cast(col, t)
        assert rc < 0 and "cast" in msg.lower(), (col, t, msg)
    rc, (v, val) = cast((np.array([3.9, -3.9, 0.0]), None), D.INT32)
    assert rc == 0 and v.tolist() == [3, -3, 0]
    rc, (v, val) = cast((np.array([5, 1 << 40], np.int64), np.array([True, False])), D.INT32)
    assert rc == 0 and v[0] == 5 and val.tolist() == [True, False]


@pytest.mark.parametrize("case", MISC["limited_batch_coalescer"]["cases"], ids=[c["name"] for c in MISC["limited_batch_coalescer"]["cases"]])
def test_gpu_filter_coalescer_reproduces_reference_batch_sizes(gpu_ctx, case):
    """coalesce/mod.rs:156-228: FilterExec pushes its filtered batches through LimitedBatchCoalescer(batch_size, fetch); with a predicate
    that keeps every row the output batch sizes must be the reference's (uint32 input batches of 8 / 100 rows, values 0..n)."""
    sizes_in = case["input_sizes"]
    assert len(set(sizes_in)) == 1
    per = sizes_in[0]
    vals = np.concatenate([np.arange(per, dtype=np.uint32) for _ in sizes_in])
    nodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.UINT32, 0, 0, 0.0), (D.EXPR_BINARY, D.OP_GTEQ, 0, 0, 0, 0.0)]
    res, sizes = gpu_filter(gpu_ctx, [(vals, None)], nodes, batch_rows=per, batch_size=case["target"], fetch=-1 if case["fetch"] is None else case["fetch"])
    assert sizes == case["expected"], case["name"]
    assert res[0][0].tolist() == vals[:sum(case["expected"])].tolist()
