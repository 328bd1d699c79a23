"""helpers shared by the Decimal128 tests: golden-vector parsing, oracle / C-ABI column and node construction"""
import re

import numpy as np

from oracle import oracle as O

_NP = {"int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64, "uint32": np.uint32, "float32": np.float32, "float64": np.float64, "bool": bool}
OPS = {"eq": O.OP_EQ, "neq": O.OP_NEQ, "lt": O.OP_LT, "lteq": O.OP_LTEQ, "gt": O.OP_GT, "gteq": O.OP_GTEQ, "plus": O.OP_PLUS, "minus": O.OP_MINUS,
       "multiply": O.OP_MULTIPLY, "divide": O.OP_DIVIDE, "modulo": O.OP_MODULO, "and": O.OP_AND, "or": O.OP_OR,
       "is_distinct_from": O.OP_IS_DISTINCT_FROM, "is_not_distinct_from": O.OP_IS_NOT_DISTINCT_FROM}


def parse_type(name):
    """'decimal128(p,s)' -> ('decimal128', p, s); 'int32' -> np.int32"""
    m = re.fullmatch(r"decimal128\((\d+),\s*(-?\d+)\)", name)
    return ("decimal128", int(m.group(1)), int(m.group(2))) if m else _NP[name]


def oracle_col(c):
    t = parse_type(c["type"])
    vals = c["values"]
    valid = np.array([v is not None for v in vals], bool)
    if isinstance(t, tuple):
        v = O.Dec([0 if x is None else x for x in vals], t[1], t[2])
    else:
        v = np.array([0 if x is None else x for x in vals], t)
    return (v, None if valid.all() else valid)


def oracle_nodes(rpn):
    nodes = []
    for item in rpn:
        if item[0] == "col":
            nodes.append((O.E_COLUMN, item[1], None, 0, 0))
        elif item[0] == "lit":
            nodes.append((O.E_LITERAL, 0, parse_type(item[1]), 1 if item[2] is None else 0, 0 if item[2] is None else item[2]))
        elif item[0] == "cast":
            nodes.append((O.E_CAST, 0, parse_type(item[1]), 0, 0))
        elif item[0] == "op":
            nodes.append((O.E_BINARY, OPS[item[1]], None, 0, 0))
        else:
            nodes.append(({"not": O.E_NOT, "is_null": O.E_IS_NULL, "is_not_null": O.E_IS_NOT_NULL, "negative": O.E_NEGATIVE}[item[0]], 0, None, 0, 0))
    return nodes


def col_as_py(col):
    """(values, valid) -> list with None for NULL; decimals as Python ints"""
    v, val = col
    out = []
    for i in range(len(v)):
        if val is not None and not val[i]:
            out.append(None)
        elif isinstance(v, O.Dec):
            out.append(int(v[i]))
        elif np.asarray(v).dtype == np.bool_:
            out.append(bool(v[i]))
        else:
            out.append(np.asarray(v)[i].item())
    return out


# ---- C ABI side ------------------------------------------------------------------------------
def gpu_type(D, t):
    if isinstance(t, tuple):
        return D.decimal128(t[1], t[2])
    return D.TYPE_OF_NP[np.dtype(t)]


def gpu_host_col(D, col, type_id=None):
    """oracle column -> HostColumn (type_id: the declared column type when numpy cannot tell, e.g. DATE32 over int32)"""
    v, val = col
    if isinstance(v, O.Dec):
        return D.HostColumn(D.decimal_to_words([int(x) for x in v]), val, D.decimal128(v.p, v.s))
    return D.HostColumn(np.asarray(v), val, type_id)


def gpu_nodes(D, rpn_or_nodes):
    """oracle node tuples -> C ABI node tuples"""
    out = []
    for kind, a, dt, is_null, lit in rpn_or_nodes:
        if kind == O.E_COLUMN:
            out.append((D.EXPR_COLUMN, a, 0, 0, 0, 0.0))
        elif kind == O.E_LITERAL:
            t = gpu_type(D, dt)
            if isinstance(dt, tuple):
                out.append((D.EXPR_LITERAL, 0, t, is_null, 0 if is_null else int(lit), 0.0))
            else:
                isf = np.dtype(dt).kind == "f"
                out.append((D.EXPR_LITERAL, 0, t, is_null, 0 if (isf or is_null) else int(lit), float(lit) if (isf and not is_null) else 0.0))
        elif kind == O.E_CAST:
            out.append((D.EXPR_CAST, 0, gpu_type(D, dt), 0, 0, 0.0))
        elif kind == O.E_BINARY:
            out.append((D.EXPR_BINARY, a, 0, 0, 0, 0.0))
        else:
            out.append((kind, 0, 0, 0, 0, 0.0))
    return out


def gpu_col_as_py(D, b, i):
    """column i of a dfgpu batch -> (list with None for NULL, type code)"""
    c = b.column(i)
    v, val = b.column_numpy(i)
    if D.type_base(c.type) == D.DECIMAL128:
        vals = D.words_to_decimal(v)
    elif c.type == D.BOOL:
        vals = [bool(x) for x in v]
    else:
        vals = [x.item() for x in v]
    return [None if (val is not None and not val[k]) else vals[k] for k in range(len(vals))], c.type


def gpu_eval(D, ctx, cols, nodes):
    """PhysicalExpr::evaluate through dfgpu_expr_evaluate_host -> (python list, type code)"""
    import ctypes as CT
    keep = [gpu_host_col(D, c) for c in cols]
    arr = (D.Column * max(len(cols), 1))(*[k.c() for k in keep])
    na = D.expr_nodes(gpu_nodes(D, nodes))
    out = CT.c_void_p()
    ctx.check(ctx.lib.dfgpu_expr_evaluate_host(ctx.h, arr, len(cols), len(cols[0][0]), na, len(nodes), CT.byref(out)))
    b = D.Batch(ctx, out.value)
    return gpu_col_as_py(D, b, 0)
