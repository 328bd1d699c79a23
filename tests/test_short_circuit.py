"""BinaryExpr short-circuit evaluation of AND / OR (binary.rs:536-600, check_short_circuit :1182-1290): which rows the RHS is evaluated on
decides where an error inside it (division by zero, failed cast) can surface.  The strategy decisions are pinned by the reference's own
test_check_short_circuit (binary.rs:5928-6110: a = [1,3,4,5,6]); the oracle restates them, the GPU interpreter must agree with the oracle
on values AND on errors."""
import numpy as np
import pytest

from oracle import oracle as O

A = np.array([1, 3, 4, 5, 6], np.int32)
Bc = np.array([1, 2, 3, 4, 5], np.int32)
I32 = np.dtype(np.int32)


def col(i): return [(O.E_COLUMN, i, None, 0, 0)]
def lit(v, dt=I32, null=False): return [(O.E_LITERAL, 0, dt, 1 if null else 0, v)]
def b(op, l, r): return l + r + [(O.E_BINARY, op, None, 0, 0)]


def div_gt0(den):           # 10 / den > 0
    return b(O.OP_GT, b(O.OP_DIVIDE, lit(10), den), lit(0))


CASES = [
    # (name, expression, expected values or None when the reference raises DivideByZero)
    ("and_all_false_returns_left", b(O.OP_AND, b(O.OP_EQ, col(0), lit(2)), div_gt0(b(O.OP_MINUS, col(0), col(0)))), [False] * 5),
    ("and_preselection_20pct_error_on_selected_row", b(O.OP_AND, b(O.OP_EQ, col(0), lit(3)), div_gt0(b(O.OP_MINUS, col(0), lit(3)))), None),
    ("and_preselection_20pct_zero_on_unselected_row", b(O.OP_AND, b(O.OP_EQ, col(0), lit(3)), div_gt0(b(O.OP_MINUS, col(0), lit(4)))), [False, False, False, False, False]),
    ("and_preselection_value", b(O.OP_AND, b(O.OP_EQ, col(0), lit(3)), div_gt0(b(O.OP_MINUS, lit(9), col(0)))), [False, True, False, False, False]),
    ("or_all_true_returns_left", b(O.OP_OR, b(O.OP_GT, col(0), lit(0)), div_gt0(b(O.OP_MINUS, col(0), col(0)))), [True] * 5),
    ("or_preselection_20pct_false", b(O.OP_OR, b(O.OP_GT, col(0), lit(2)), div_gt0(b(O.OP_MINUS, col(0), lit(3)))), [False, True, True, True, True]),
    ("or_60pct_false_full_evaluation_raises", b(O.OP_OR, b(O.OP_GT, col(0), lit(4)), div_gt0(b(O.OP_MINUS, col(0), lit(4)))), None),
    ("and_60pct_true_full_evaluation_raises", b(O.OP_AND, b(O.OP_GT, col(0), lit(3)), div_gt0(b(O.OP_MINUS, col(0), lit(1)))), None),
    ("scalar_false_and_returns_left", b(O.OP_AND, lit(False, np.dtype(bool)), div_gt0(b(O.OP_MINUS, col(0), col(0)))), [False] * 5),
    ("scalar_true_or_returns_left", b(O.OP_OR, lit(True, np.dtype(bool)), div_gt0(b(O.OP_MINUS, col(0), col(0)))), [True] * 5),
    ("scalar_true_and_returns_right", b(O.OP_AND, lit(True, np.dtype(bool)), div_gt0(b(O.OP_MINUS, col(0), col(0)))), None),
]


@pytest.mark.parametrize("name,expr,expected", CASES, ids=[c[0] for c in CASES])
def test_oracle_short_circuit(name, expr, expected):
    cols = [(A, None), (Bc, None)]
    if expected is None:
        with pytest.raises(O.ArrowDivideByZero):
            O.eval_expr(cols, expr)
    else:
        v, val = O.eval_expr(cols, expr)
        assert val is None or np.asarray(val).all()
        assert np.asarray(v, bool).tolist() == expected


def test_oracle_lhs_with_nulls_is_not_short_circuited():
    # c = [T, F, NULL, T, NULL] (binary.rs:6010-6060): "Mixed values with nulls - shouldn't short-circuit" -> the RHS runs on every row
    c = (np.array([True, False, False, True, False]), np.array([True, True, False, True, False]))
    x = (np.array([1, 0, 1, 1, 1], np.int32), None)
    expr = b(O.OP_AND, col(0), div_gt0(col(1)))
    with pytest.raises(O.ArrowDivideByZero):          # row 1 has LHS false and x = 0: still evaluated, still an error
        O.eval_expr([c, x], expr)
    c2 = (np.array([True, False, False, True, False]), None)   # no NULLs, 40 % true -> neither skip nor pre-selection: full evaluation
    with pytest.raises(O.ArrowDivideByZero):
        O.eval_expr([c2, x], expr)
    c3 = (np.array([False, False, False, True, False]), None)  # 20 % true -> pre-selection: row 1 is never evaluated
    v, val = O.eval_expr([c3, x], expr)
    assert np.asarray(v, bool).tolist() == [False, False, False, True, False]


@pytest.mark.gpu
@pytest.mark.parametrize("name,expr,expected", CASES, ids=[c[0] for c in CASES])
def test_gpu_short_circuit_matches_oracle(gpu_ctx, name, expr, expected):
    from datafusion_b200 import capi as D
    from test_gpu_filter import NP2T
    import ctypes as CT
    nodes = []
    for kind, a, dt, is_null, v in expr:
        if kind == O.E_LITERAL:
            nodes.append((D.EXPR_LITERAL, 0, NP2T[np.dtype(dt)], is_null, int(v), 0.0))
        else:
            nodes.append((kind, a, 0, 0, 0, 0.0))
    keep = [D.HostColumn(A), D.HostColumn(Bc)]
    arr = (D.Column * 2)(*[k.c() for k in keep])
    out = CT.c_void_p()
    rc = gpu_ctx.lib.dfgpu_expr_evaluate_host(gpu_ctx.h, arr, 2, 5, D.expr_nodes(nodes), len(nodes), CT.byref(out))
    if expected is None:
        assert rc == -4, f"expected DivideByZero, got rc {rc}"
        return
    gpu_ctx.check(rc)
    v, val = D.Batch(gpu_ctx, out.value).column_numpy(0)
    assert (val is None or val.all()) and v.tolist() == expected
    # the same predicate through FilterExec (fused interpreter path) keeps exactly the TRUE rows
    from harness import gpu_filter
    got, _ = gpu_filter(gpu_ctx, [(A, None), (Bc, None)], nodes)
    assert got[0][0].tolist() == A[np.array(expected)].tolist()
