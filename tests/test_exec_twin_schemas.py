"""Host logic of the operator twin (datafusion_b200/exec.py) pinned by the reference's own schema rules — no GPU needed.
  * build_join_schema: the 16 cases of `test_join_schema` (physical-plan/src/joins/utils.rs:2827-2879) + the semi / anti / mark shapes of
    `build_join_schema` (:266-345);
  * Sum::return_type (functions-aggregate/src/sum.rs:232-261) and the Partial state schemas the Final modes consume
    (aggregates/mod.rs:3591-3700 snapshots: AVG state = [count, sum])."""
import pyarrow as pa
import pytest

from datafusion_b200.exec import AggregateExpr, GpuAggregateExec, MemoryExec, build_join_schema

A, A_NULLS = pa.schema([pa.field("a", pa.int32(), False)]), pa.schema([pa.field("a", pa.int32(), True)])
B, B_NULLS = pa.schema([pa.field("b", pa.int32(), False)]), pa.schema([pa.field("b", pa.int32(), True)])

CASES = [  # (left_in, right_in, join_type, left_out, right_out) — utils.rs:2833-2853, verbatim
    (A, B, "Inner", A, B), (A, B_NULLS, "Inner", A, B_NULLS), (A_NULLS, B, "Inner", A_NULLS, B), (A_NULLS, B_NULLS, "Inner", A_NULLS, B_NULLS),
    (A, B, "Left", A, B_NULLS), (A, B_NULLS, "Left", A, B_NULLS), (A_NULLS, B, "Left", A_NULLS, B_NULLS), (A_NULLS, B_NULLS, "Left", A_NULLS, B_NULLS),
    (A, B, "Right", A_NULLS, B), (A, B_NULLS, "Right", A_NULLS, B_NULLS), (A_NULLS, B, "Right", A_NULLS, B), (A_NULLS, B_NULLS, "Right", A_NULLS, B_NULLS),
    (A, B, "Full", A_NULLS, B_NULLS), (A, B_NULLS, "Full", A_NULLS, B_NULLS), (A_NULLS, B, "Full", A_NULLS, B_NULLS), (A_NULLS, B_NULLS, "Full", A_NULLS, B_NULLS),
]


@pytest.mark.parametrize("left_in,right_in,join_type,left_out,right_out", CASES)
def test_join_schema(left_in, right_in, join_type, left_out, right_out):
    schema, idx = build_join_schema(left_in, right_in, join_type)
    assert schema == pa.schema(list(left_out) + list(right_out)), (join_type, left_in, right_in)
    assert idx == [(0, 0), (1, 0)]


def test_join_schema_semi_anti_mark():
    l = pa.schema([pa.field("a", pa.int32(), False), pa.field("x", pa.int64(), True)])
    r = pa.schema([pa.field("b", pa.int32(), True)])
    for jt in ("LeftSemi", "LeftAnti"):
        assert build_join_schema(l, r, jt) == (l, [(0, 0), (0, 1)])
    for jt in ("RightSemi", "RightAnti"):
        assert build_join_schema(l, r, jt) == (r, [(1, 0)])
    mark = pa.field("mark", pa.bool_(), False)                       # Field::new("mark", DataType::Boolean, false), JoinSide::None
    assert build_join_schema(l, r, "LeftMark") == (pa.schema(list(l) + [mark]), [(0, 0), (0, 1), (2, 0)])
    assert build_join_schema(l, r, "RightMark") == (pa.schema(list(r) + [mark]), [(1, 0), (2, 0)])
    with pytest.raises(ValueError):
        build_join_schema(l, r, "Cross")


def test_sum_return_type_rule():
    e = AggregateExpr("sum", "v")
    assert e.value_type(pa.int8()) == e.value_type(pa.int32()) == e.value_type(pa.int64()) == pa.int64()      # signed -> Int64
    assert e.value_type(pa.uint16()) == e.value_type(pa.uint64()) == pa.uint64()                              # unsigned -> UInt64
    assert e.value_type(pa.float32()) == e.value_type(pa.float64()) == pa.float64()                           # floats -> Float64
    assert e.value_type(pa.decimal128(15, 2)) == pa.decimal128(25, 2)                                          # precision + 10, same scale
    assert e.value_type(pa.decimal128(35, 4)) == pa.decimal128(38, 4)                                          # capped at DECIMAL128_MAX_PRECISION
    assert AggregateExpr("count", "v").value_type(pa.float64()) == pa.int64() and AggregateExpr("avg", "v").value_type(pa.int32()) == pa.float64()
    assert AggregateExpr("min", "v").value_type(pa.date32()) == pa.date32()


def test_partial_state_schema_and_final_output_schema():
    t = pa.table({"a": pa.array([2, 3], pa.uint32()), "b": pa.array([1.0, 2.0], pa.float64())})
    src = MemoryExec(t.to_batches(), t.schema)
    part = GpuAggregateExec("Partial", ["a"], [AggregateExpr("avg", "b", "AVG(b)"), AggregateExpr("sum", "b", "s"), AggregateExpr("count", "b", "n")], src)
    assert [(f.name, f.type) for f in part.schema] == [("a", pa.uint32()), ("AVG(b)[count]", pa.uint64()), ("AVG(b)[sum]", pa.float64()),
                                                       ("s[sum]", pa.float64()), ("n[count]", pa.int64())]
    fin = GpuAggregateExec("Final", ["a"], [AggregateExpr("avg", "b", "AVG(b)"), AggregateExpr("sum", "b", "s"), AggregateExpr("count", "b", "n")], part, input_schema=t.schema)
    assert [(f.name, f.type) for f in fin.schema] == [("a", pa.uint32()), ("AVG(b)", pa.float64()), ("s", pa.float64()), ("n", pa.int64())]
    assert fin.state_input and not fin.state_output and part.state_output and not part.state_input


# ---- result types of decimal expressions: the twin's BinaryExpr.data_type against the types the reference's own decimal tests assert
# (binary.rs:4355-5000, transcribed in tests/golden/decimal_kat.json; arrow-arith's decimal_op rules) ----
import json
import os
import re

from datafusion_b200 import capi as D
from datafusion_b200.exec import BinaryExpr, CastExpr, Column, Literal

_OPS = {"plus": D.OP_PLUS, "minus": D.OP_MINUS, "multiply": D.OP_MULTIPLY, "divide": D.OP_DIVIDE, "modulo": D.OP_MODULO, "eq": D.OP_EQ, "neq": D.OP_NEQ,
        "lt": D.OP_LT, "lteq": D.OP_LTEQ, "gt": D.OP_GT, "gteq": D.OP_GTEQ, "is_distinct_from": D.OP_IS_DISTINCT_FROM, "is_not_distinct_from": D.OP_IS_NOT_DISTINCT_FROM}


def _pa_type(s):
    m = re.fullmatch(r"decimal128\((\d+),(-?\d+)\)", s)
    if m:
        return pa.decimal128(int(m.group(1)), int(m.group(2)))
    return {"bool": pa.bool_(), "int32": pa.int32(), "int64": pa.int64(), "float64": pa.float64(), "float32": pa.float32()}[s]


def _decimal_cases():
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "decimal_kat.json")))
    return [c for c in d["cases"] if "expected" in c and all(it[0] in ("col", "lit", "cast", "op") for it in c["rpn"]) and all(it[0] != "op" or it[1] in _OPS for it in c["rpn"])]


@pytest.mark.parametrize("case", _decimal_cases(), ids=lambda c: c["name"])
def test_decimal_expression_result_types_match_the_reference_vectors(case):
    schema = pa.schema([pa.field(f"c{i}", _pa_type(c["type"])) for i, c in enumerate(case["cols"])])
    st = []
    for it in case["rpn"]:
        if it[0] == "col":
            st.append(Column(f"c{it[1]}"))
        elif it[0] == "lit":
            st.append(Literal(it[2], _pa_type(it[1])))
        elif it[0] == "cast":
            st.append(CastExpr(st.pop(), _pa_type(it[1])))
        else:
            r, l = st.pop(), st.pop()
            st.append(BinaryExpr(l, _OPS[it[1]], r))
    assert len(st) == 1
    assert st[0].data_type(schema) == _pa_type(case["expected"]["type"]), case["name"]
    out = []
    st[0].rpn(schema, out)                                  # lowers without error, one node per RPN item
    assert len(out) == len(case["rpn"])
