"""Host-side mirror of the reference's operator interface for the hot path.

The reference's operators implement `trait ExecutionPlan` (datafusion/physical-plan/src/execution_plan.rs:102;
`execute(partition, ctx) -> SendableRecordBatchStream` :696) and are driven by `collect(plan, ctx)` (:1752).
No Rust toolchain exists in this image, so this module is the Python stand-in for the Rust shim
(`GpuFilterExec` / `GpuHashJoinExec` / `GpuAggregateExec`): same constructor
arguments, same output schemas, same error behaviour — all compute goes through the C ABI of
libdfgpu.so (capi.py); nothing here computes on the CPU.

Data model = pyarrow RecordBatch (crossing the boundary as Arrow C Data Interface structs, exactly as
datafusion/ffi/src/record_batch_stream.rs:101-167 does).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import pyarrow as pa

from . import capi as D


# ---------------------------------------------------------------------------------------------
# types
# ---------------------------------------------------------------------------------------------
def type_id(t: pa.DataType) -> int:
    m = {pa.bool_(): D.BOOL, pa.int8(): D.INT8, pa.int16(): D.INT16, pa.int32(): D.INT32, pa.int64(): D.INT64,
         pa.uint8(): D.UINT8, pa.uint16(): D.UINT16, pa.uint32(): D.UINT32, pa.uint64(): D.UINT64,
         pa.float32(): D.FLOAT32, pa.float64(): D.FLOAT64, pa.date32(): D.DATE32, pa.date64(): D.DATE64}
    if t in m:
        return m[t]
    if pa.types.is_timestamp(t):
        return D.TIMESTAMP
    if pa.types.is_decimal128(t):
        return D.decimal128(t.precision, t.scale) if 0 <= t.scale <= t.precision else D.DECIMAL128
    raise NotImplementedError(f"This feature is not implemented: GPU operators do not support Arrow type {t}")


def arrow_type(tid: int) -> pa.DataType:
    if D.type_base(tid) == D.DECIMAL128:
        p, sc = D.decimal_precision_scale(tid)
        return pa.decimal128(p or 38, sc)
    return {D.BOOL: pa.bool_(), D.INT8: pa.int8(), D.INT16: pa.int16(), D.INT32: pa.int32(), D.INT64: pa.int64(), D.UINT8: pa.uint8(),
            D.UINT16: pa.uint16(), D.UINT32: pa.uint32(), D.UINT64: pa.uint64(), D.FLOAT32: pa.float32(), D.FLOAT64: pa.float64(),
            D.DATE32: pa.date32(), D.DATE64: pa.date64(), D.TIMESTAMP: pa.timestamp("ns")}[tid]


# ---------------------------------------------------------------------------------------------
# expressions — PhysicalExpr (physical-expr-common/src/physical_expr.rs:76)
# ---------------------------------------------------------------------------------------------
class Expr:
    def _bin(self, op, other):
        return BinaryExpr(self, op, other if isinstance(other, Expr) else lit(other))

    def __gt__(self, o): return self._bin(D.OP_GT, o)
    def __ge__(self, o): return self._bin(D.OP_GTEQ, o)
    def __lt__(self, o): return self._bin(D.OP_LT, o)
    def __le__(self, o): return self._bin(D.OP_LTEQ, o)
    def __eq__(self, o): return self._bin(D.OP_EQ, o)  # type: ignore[override]
    def __ne__(self, o): return self._bin(D.OP_NEQ, o)  # type: ignore[override]
    def __add__(self, o): return self._bin(D.OP_PLUS, o)
    def __sub__(self, o): return self._bin(D.OP_MINUS, o)
    def __mul__(self, o): return self._bin(D.OP_MULTIPLY, o)
    def __truediv__(self, o): return self._bin(D.OP_DIVIDE, o)
    def __mod__(self, o): return self._bin(D.OP_MODULO, o)
    def __and__(self, o): return self._bin(D.OP_AND, o)
    def __or__(self, o): return self._bin(D.OP_OR, o)
    def __invert__(self): return UnaryExpr(D.EXPR_NOT, self)
    def __neg__(self): return UnaryExpr(D.EXPR_NEGATIVE, self)
    def is_null(self): return UnaryExpr(D.EXPR_IS_NULL, self)
    def is_not_null(self): return UnaryExpr(D.EXPR_IS_NOT_NULL, self)
    def is_distinct_from(self, o): return self._bin(D.OP_IS_DISTINCT_FROM, o)
    def is_not_distinct_from(self, o): return self._bin(D.OP_IS_NOT_DISTINCT_FROM, o)
    def cast(self, t: pa.DataType): return CastExpr(self, t)
    __hash__ = None  # type: ignore[assignment]

    def data_type(self, schema: pa.Schema) -> pa.DataType:
        raise NotImplementedError

    def rpn(self, schema: pa.Schema, out: list) -> None:
        raise NotImplementedError


class Column(Expr):
    """expressions/column.rs:121"""

    def __init__(self, name: str):
        self.name = name

    def data_type(self, schema): return schema.field(self.name).type

    def rpn(self, schema, out):
        ix = schema.get_field_index(self.name)
        if ix < 0:
            raise KeyError(f"Schema error: No field named {self.name}")
        out.append((D.EXPR_COLUMN, ix, 0, 0, 0, 0.0))


class Literal(Expr):
    """expressions/literal.rs:106; value None = NULL of the given type"""

    def __init__(self, value, type: Optional[pa.DataType] = None):
        if type is None:
            type = pa.bool_() if isinstance(value, bool) else pa.int64() if isinstance(value, int) else pa.float64()
        self.value, self.type = value, type

    def data_type(self, schema): return self.type

    def rpn(self, schema, out):
        tid = type_id(self.type)
        isnull = 1 if self.value is None else 0
        v = 0 if self.value is None else self.value
        if tid in (D.FLOAT32, D.FLOAT64):
            out.append((D.EXPR_LITERAL, 0, tid, isnull, 0, float(v)))
        elif D.type_base(tid) == D.DECIMAL128:
            import decimal
            with decimal.localcontext() as dctx:
                dctx.prec = 80
                unscaled = int(decimal.Decimal(str(v)).scaleb(self.type.scale).to_integral_value(rounding=decimal.ROUND_HALF_UP))
            out.append((D.EXPR_LITERAL, 0, tid, isnull, unscaled, 0.0))   # capi.expr_nodes splits the 128-bit value
        else:
            if hasattr(v, "toordinal") and tid == D.DATE32:
                import datetime
                v = (v - datetime.date(1970, 1, 1)).days
            out.append((D.EXPR_LITERAL, 0, tid, isnull, int(v), 0.0))


class BinaryExpr(Expr):
    """expressions/binary.rs:536-676.  Operand types must already agree (the planner's type coercion);
    as a convenience an untyped Python literal adopts the other side's type."""

    def __init__(self, left: Expr, op: int, right: Expr):
        self.left, self.op, self.right = left, op, right

    def _coerced(self, schema):
        l, r = self.left, self.right
        if isinstance(r, Literal) and not isinstance(l, Literal):
            lt = l.data_type(schema)
            if r.type != lt and (pa.types.is_integer(r.type) or pa.types.is_floating(r.type)) and not pa.types.is_boolean(lt):
                if pa.types.is_decimal128(lt) and self.op in (D.OP_PLUS, D.OP_MINUS, D.OP_MULTIPLY, D.OP_DIVIDE, D.OP_MODULO) and pa.types.is_integer(r.type):
                    r = Literal(r.value, pa.decimal128(20, 0))    # Int64 -> Decimal128(20, 0) (type_coercion/binary.rs:1265)
                else:
                    r = Literal(r.value, lt)
        elif isinstance(l, Literal) and not isinstance(r, Literal):
            rt = r.data_type(schema)
            if l.type != rt and (pa.types.is_integer(l.type) or pa.types.is_floating(l.type)) and not pa.types.is_boolean(rt):
                if pa.types.is_decimal128(rt) and self.op in (D.OP_PLUS, D.OP_MINUS, D.OP_MULTIPLY, D.OP_DIVIDE, D.OP_MODULO) and pa.types.is_integer(l.type):
                    l = Literal(l.value, pa.decimal128(20, 0))
                else:
                    l = Literal(l.value, rt)
        return l, r

    def data_type(self, schema):
        if self.op in (D.OP_EQ, D.OP_NEQ, D.OP_LT, D.OP_LTEQ, D.OP_GT, D.OP_GTEQ, D.OP_AND, D.OP_OR, D.OP_IS_DISTINCT_FROM,
                       D.OP_IS_NOT_DISTINCT_FROM):
            return pa.bool_()
        l, r = self._coerced(schema)
        lt, rt = l.data_type(schema), r.data_type(schema)
        if pa.types.is_decimal128(lt) and pa.types.is_decimal128(rt):   # arrow-arith decimal_op result types (include/dfgpu.h)
            p1, s1, p2, s2 = lt.precision, lt.scale, rt.precision, rt.scale
            if self.op in (D.OP_PLUS, D.OP_MINUS):
                sc = max(s1, s2); return pa.decimal128(min(38, sc + max(p1 - s1, p2 - s2) + 1), sc)
            if self.op == D.OP_MULTIPLY:
                return pa.decimal128(min(38, p1 + p2 + 1), s1 + s2)
            if self.op == D.OP_DIVIDE:
                sc = min(38, s1 + 4); return pa.decimal128(min(38, sc - s1 + s2 + p1), sc)
            if self.op == D.OP_MODULO:
                sc = max(s1, s2); return pa.decimal128(min(38, sc + min(p1 - s1, p2 - s2)), sc)
        return lt

    def rpn(self, schema, out):
        l, r = self._coerced(schema)
        l.rpn(schema, out)
        r.rpn(schema, out)
        out.append((D.EXPR_BINARY, self.op, 0, 0, 0, 0.0))


class UnaryExpr(Expr):
    def __init__(self, kind: int, arg: Expr):
        self.kind, self.arg = kind, arg

    def data_type(self, schema):
        return self.arg.data_type(schema) if self.kind == D.EXPR_NEGATIVE else pa.bool_()

    def rpn(self, schema, out):
        self.arg.rpn(schema, out)
        out.append((self.kind, 0, 0, 0, 0, 0.0))


class CastExpr(Expr):
    def __init__(self, arg: Expr, to: pa.DataType):
        self.arg, self.to = arg, to

    def data_type(self, schema): return self.to

    def rpn(self, schema, out):
        self.arg.rpn(schema, out)
        out.append((D.EXPR_CAST, 0, type_id(self.to), 0, 0, 0.0))


def col(name: str) -> Column: return Column(name)
def lit(value, type: Optional[pa.DataType] = None) -> Literal: return Literal(value, type)


# ---------------------------------------------------------------------------------------------
# plans
# ---------------------------------------------------------------------------------------------
class SessionConfig:
    """the execution.* keys that change hot-path behaviour (common/src/config.rs:904-923)"""

    def __init__(self, batch_size: int = 8192, perfect_hash_join_small_build_threshold: int = 1024,
                 perfect_hash_join_min_key_density: float = 0.15, force_hash_collisions: bool = False, device: int = 0):
        self.batch_size = batch_size
        self.perfect_hash_join_small_build_threshold = perfect_hash_join_small_build_threshold
        self.perfect_hash_join_min_key_density = perfect_hash_join_min_key_density
        self.force_hash_collisions = force_hash_collisions
        self.device = device


class TaskContext:
    """execution/src/task.rs:52 — owns the dfgpu context (device + stream)"""

    def __init__(self, config: Optional[SessionConfig] = None, ctx: Optional[D.Context] = None):
        self.config = config or SessionConfig()
        self.gpu = ctx or D.Context(self.config.device)


class ExecutionPlan:
    schema: pa.Schema

    def children(self) -> List["ExecutionPlan"]: return []
    def name(self) -> str: return type(self).__name__
    def execute(self, ctx: TaskContext) -> Iterator[pa.RecordBatch]: raise NotImplementedError
    def metrics(self) -> dict: return {}


class MemoryExec(ExecutionPlan):
    """TestMemoryExec (physical-plan/src/test.rs): yields the given batches of ONE partition"""

    def __init__(self, batches: Sequence[pa.RecordBatch], schema: Optional[pa.Schema] = None):
        self.batches = list(batches)
        self.schema = schema or self.batches[0].schema

    def execute(self, ctx):
        return iter(self.batches)


def _rename(rb: pa.RecordBatch, schema: pa.Schema) -> pa.RecordBatch:
    cols = []
    for c, f in zip(rb.columns, schema):
        cols.append(c if c.type == f.type else c.cast(f.type))
    return pa.RecordBatch.from_arrays(cols, schema=schema)


def _drain(op, schema) -> Iterator[pa.RecordBatch]:
    while True:
        b = op.next(host=True)
        if b is None:
            return
        yield _rename(b.to_arrow(), schema)


# ---------------------------------------------------------------------------------------------
# string keys — Utf8 / Utf8View / Dictionary(_, Utf8) columns as INT32 codes in ONE code space (dfgpu_dictionary)
# ---------------------------------------------------------------------------------------------
def _is_string_like(t: pa.DataType) -> bool:
    if pa.types.is_dictionary(t):
        t = t.value_type
    return pa.types.is_string(t) or pa.types.is_large_string(t) or (hasattr(pa.types, "is_string_view") and pa.types.is_string_view(t))


class StringDictionary:
    """the plan-wide code space: equal strings <=> equal codes in every column, every batch and on both sides of a join"""

    def __init__(self, ctx: TaskContext):
        self.ctx = ctx
        self.dic = D.Dictionary(ctx.gpu)

    def encode(self, arr: pa.Array) -> pa.Array:
        """string-like array -> int32 codes (NULL stays NULL): the batch's own dictionary is unified on the host (distinct values only),
        the row codes are rewritten on the device (dfgpu_dictionary_remap)"""
        import numpy as np
        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        d = arr if pa.types.is_dictionary(arr.type) else arr.dictionary_encode()
        values = d.dictionary.cast(pa.string())
        bufs = values.buffers()
        offsets = np.frombuffer(bufs[1], np.int32, len(values) + 1, values.offset * 4) if len(values) else np.zeros(1, np.int32)
        data = np.frombuffer(bufs[2], np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
        vvalid = None if values.null_count == 0 else np.asarray(values.is_valid())
        remap = self.dic.unify(offsets, data, vvalid)
        idx = d.indices.cast(pa.int32())
        codes = D.HostColumn(np.asarray(idx.fill_null(0)), None if idx.null_count == 0 else np.asarray(idx.is_valid()))
        b = self.dic.remap(codes, remap, on_host=True)
        v, valid = b.column_numpy(0)
        b.release()
        return pa.array(v, pa.int32(), mask=None if valid is None else ~valid)

    def decode(self, codes: pa.Array, to: pa.DataType) -> pa.Array:
        n = self.dic.size()
        values = pa.array([self.dic.value(i).decode() for i in range(n)], pa.string())
        if isinstance(codes, pa.ChunkedArray):
            codes = codes.combine_chunks()
        out = pa.DictionaryArray.from_arrays(codes.cast(pa.int32()), values)
        return out if pa.types.is_dictionary(to) else out.cast(to)

    def code(self, s: str) -> int:
        """the literal of `col = 'text'`: -1 when the string was never seen (matches nothing)"""
        return self.dic.code(s.encode())

    def close(self):
        self.dic.close()


class DictionaryEncodeExec(ExecutionPlan):
    """string-like columns of the input -> INT32 code columns (same names); everything else passes through"""

    def __init__(self, input: ExecutionPlan, dictionary_of: "callable"):
        self.input, self.dictionary_of = input, dictionary_of
        self.string_cols = [i for i, f in enumerate(input.schema) if _is_string_like(f.type)]
        self.schema = pa.schema([pa.field(f.name, pa.int32(), True) if i in self.string_cols else f for i, f in enumerate(input.schema)])

    def children(self): return [self.input]

    def execute(self, ctx):
        sd = self.dictionary_of(ctx)
        for rb in self.input.execute(ctx):
            cols = [sd.encode(rb.column(i)) if i in self.string_cols else rb.column(i) for i in range(rb.num_columns)]
            yield pa.RecordBatch.from_arrays(cols, schema=self.schema)


class DictionaryDecodeExec(ExecutionPlan):
    """INT32 code columns `names` of the input -> strings of type `to` (after the GPU operators)"""

    def __init__(self, input: ExecutionPlan, names: Sequence[str], dictionary_of: "callable", to: pa.DataType = pa.string()):
        self.input, self.names, self.dictionary_of, self.to = input, list(names), dictionary_of, to
        self.schema = pa.schema([pa.field(f.name, to, True) if f.name in self.names else f for f in input.schema])

    def children(self): return [self.input]

    def execute(self, ctx):
        sd = self.dictionary_of(ctx)
        for rb in self.input.execute(ctx):
            cols = [sd.decode(rb.column(i), self.to) if f.name in self.names else rb.column(i) for i, f in enumerate(rb.schema)]
            yield pa.RecordBatch.from_arrays(cols, schema=self.schema)


def plan_string_dictionary():
    """a factory for the `dictionary_of` argument: one StringDictionary per TaskContext, created on first use"""
    cache = {}

    def get(ctx: TaskContext) -> StringDictionary:
        if id(ctx) not in cache:
            cache[id(ctx)] = StringDictionary(ctx)
        return cache[id(ctx)]
    return get


class GpuFilterExec(ExecutionPlan):
    """FilterExec (physical-plan/src/filter.rs:85): FilterExecBuilder::new(predicate, input).with_projection(..).with_fetch(..)"""

    def __init__(self, predicate: Expr, input: ExecutionPlan, projection: Optional[Sequence[int]] = None, fetch: Optional[int] = None):
        self.predicate, self.input, self.projection, self.fetch = predicate, input, projection, fetch
        if predicate.data_type(input.schema) != pa.bool_():
            # filter.rs:139-145
            raise ValueError(f"Error during planning: Filter predicate must return BOOLEAN values, got {predicate.data_type(input.schema)}")
        fields = list(input.schema) if projection is None else [input.schema.field(i) for i in projection]
        self.schema = pa.schema(fields)
        self._metrics = {}

    def children(self): return [self.input]

    def execute(self, ctx):
        nodes: list = []
        self.predicate.rpn(self.input.schema, nodes)
        types = [type_id(f.type) for f in self.input.schema]
        op = D.FilterHandle(ctx.gpu, types, nodes, self.projection, ctx.config.batch_size, -1 if self.fetch is None else self.fetch)
        try:
            for rb in self.input.execute(ctx):
                op.push_arrow(rb)
                yield from _drain(op, self.schema)
            op.finish()
            yield from _drain(op, self.schema)
            self._metrics = {k: op.metric(k) for k in ("input_rows", "output_rows", "selectivity_num", "selectivity_den")}
        finally:
            op.close()

    def metrics(self): return self._metrics


class GpuProjectionExec(ExecutionPlan):
    """ProjectionExec over PhysicalExpr::evaluate (physical-expr-common/src/physical_expr.rs:88): each output column is
    one expression evaluated on the GPU; plain `Column` expressions are passed through untouched (zero copy on the host)."""

    def __init__(self, exprs: Sequence[Tuple[Expr, str]], input: ExecutionPlan):
        self.exprs, self.input = list(exprs), input
        self.schema = pa.schema([pa.field(name, e.data_type(input.schema)) for e, name in self.exprs])

    def children(self): return [self.input]

    def execute(self, ctx):
        import ctypes as C
        import numpy as np
        isch = self.input.schema
        for rb in self.input.execute(ctx):
            cols = []
            for e, name in self.exprs:
                if isinstance(e, Column):
                    cols.append(rb.column(isch.get_field_index(e.name)))
                    continue
                nodes: list = []
                e.rpn(isch, nodes)
                hcols, keep = [], []
                for i, f in enumerate(isch):
                    arr = rb.column(i)
                    if pa.types.is_date32(f.type):
                        vals = np.asarray(arr.cast(pa.int32()).fill_null(0))
                    elif pa.types.is_boolean(f.type):
                        vals = np.asarray(arr.fill_null(False))
                    else:
                        vals = np.asarray(arr.fill_null(0))
                    valid = None if arr.null_count == 0 else ~np.asarray(arr.is_null())
                    keep.append(D.HostColumn(vals, valid, type_id(f.type)))
                arrc = (D.Column * len(keep))(*[k.c() for k in keep])
                na = D.expr_nodes(nodes)
                out = C.c_void_p()
                ctx.gpu.check(ctx.gpu.lib.dfgpu_expr_evaluate_host(ctx.gpu.h, arrc, len(keep), rb.num_rows, na, len(nodes), C.byref(out)))
                b = D.Batch(ctx.gpu, out.value)
                res = b.to_arrow().column(0)
                t = e.data_type(isch)
                cols.append(res if res.type == t else res.cast(t))
            yield pa.RecordBatch.from_arrays(cols, schema=self.schema)


_JOIN_TYPES = {"Inner": D.JOIN_INNER, "Left": D.JOIN_LEFT, "Right": D.JOIN_RIGHT, "Full": D.JOIN_FULL, "LeftSemi": D.JOIN_LEFT_SEMI,
               "RightSemi": D.JOIN_RIGHT_SEMI, "LeftAnti": D.JOIN_LEFT_ANTI, "RightAnti": D.JOIN_RIGHT_ANTI, "LeftMark": D.JOIN_LEFT_MARK,
               "RightMark": D.JOIN_RIGHT_MARK}


def build_join_schema(left: pa.Schema, right: pa.Schema, join_type: str) -> Tuple[pa.Schema, List[Tuple[int, int]]]:
    """joins/utils.rs build_join_schema: output fields + ColumnIndex (side, index); side 0=left 1=right 2=mark"""
    def nullable(fields): return [pa.field(f.name, f.type, True) for f in fields]
    l, r = list(left), list(right)
    if join_type in ("Inner", "Left", "Right", "Full"):
        lf = nullable(l) if join_type in ("Right", "Full") else l
        rf = nullable(r) if join_type in ("Left", "Full") else r
        return pa.schema(lf + rf), [(0, i) for i in range(len(l))] + [(1, i) for i in range(len(r))]
    if join_type in ("LeftSemi", "LeftAnti"):
        return pa.schema(l), [(0, i) for i in range(len(l))]
    if join_type in ("RightSemi", "RightAnti"):
        return pa.schema(r), [(1, i) for i in range(len(r))]
    if join_type == "LeftMark":
        return pa.schema(l + [pa.field("mark", pa.bool_(), False)]), [(0, i) for i in range(len(l))] + [(2, 0)]
    if join_type == "RightMark":
        return pa.schema(r + [pa.field("mark", pa.bool_(), False)]), [(1, i) for i in range(len(r))] + [(2, 0)]
    raise ValueError(join_type)


class JoinFilter:
    """joins/utils.rs JoinFilter: `expression` over an intermediate batch whose column c ("f0", "f1", ...) is
    column_indices[c] = (side "left"|"right", index)"""

    def __init__(self, expression: Expr, column_indices: Sequence[Tuple[str, int]]):
        self.expression, self.column_indices = expression, list(column_indices)


class GpuHashJoinExec(ExecutionPlan):
    """HashJoinExec::try_new(left, right, on, filter, join_type, projection, partition_mode, null_equality, null_aware)
    (physical-plan/src/joins/hash_join/exec.rs:752).  left = build side, right = probe side."""

    def __init__(self, left: ExecutionPlan, right: ExecutionPlan, on: Sequence[Tuple[str, str]], join_type: str = "Inner",
                 null_equality: str = "NullEqualsNothing", filter=None, projection: Optional[Sequence[int]] = None, null_aware: bool = False):
        if not on:
            raise ValueError("Error during planning: On constraints in HashJoinExec should be non-empty")  # exec.rs try_new
        self.filter, self.null_aware = filter, bool(null_aware)
        self.left, self.right, self.on, self.join_type, self.null_equality = left, right, list(on), join_type, null_equality
        full, idx = build_join_schema(left.schema, right.schema, join_type)
        if projection is not None:
            full = pa.schema([full.field(i) for i in projection])
            idx = [idx[i] for i in projection]
        self.schema, self.column_indices = full, idx
        self._metrics = {}

    def children(self): return [self.left, self.right]

    def execute(self, ctx):
        cfg = ctx.config
        bt = [type_id(f.type) for f in self.left.schema]
        pt = [type_id(f.type) for f in self.right.schema]
        ob = [self.left.schema.get_field_index(l) for l, _ in self.on]
        op_ = [self.right.schema.get_field_index(r) for _, r in self.on]
        if min(ob + op_) < 0:
            raise KeyError("Schema error: join key not found")
        op = D.HashJoinHandle(ctx.gpu, bt, pt, ob, op_, [s for s, _ in self.column_indices], [i for _, i in self.column_indices],
                              _JOIN_TYPES[self.join_type], D.NULL_EQUALS_NULL if self.null_equality == "NullEqualsNull" else D.NULL_EQUALS_NOTHING,
                              cfg.batch_size, cfg.perfect_hash_join_small_build_threshold, cfg.perfect_hash_join_min_key_density,
                              cfg.force_hash_collisions, self.null_aware)
        if self.filter is not None:
            fields = [(self.left.schema if sd == "left" else self.right.schema).field(ix) for sd, ix in self.filter.column_indices]
            inter = pa.schema([pa.field(f"f{i}", f.type) for i, f in enumerate(fields)])
            nodes: list = []
            self.filter.expression.rpn(inter, nodes)
            op.set_filter([0 if sd == "left" else 1 for sd, _ in self.filter.column_indices], [ix for _, ix in self.filter.column_indices], nodes)
        try:
            for rb in self.left.execute(ctx):     # collect_left_input
                op.push_build_arrow(rb)
            op.finish_build()
            for rb in self.right.execute(ctx):    # FetchProbeBatch / ProcessProbeBatch
                op.push_probe_arrow(rb)
                yield from _drain(op, self.schema)
            op.finish_probe()                     # ExhaustedProbeSide
            yield from _drain(op, self.schema)
            self._metrics = {k: op.metric(k) for k in ("build_input_rows", "input_rows", "output_rows", "array_map_created_count", "probe_hits")}
        finally:
            op.close()

    def metrics(self): return self._metrics


_AGG_FUNCS = {"sum": D.AGG_SUM, "count": D.AGG_COUNT, "min": D.AGG_MIN, "max": D.AGG_MAX, "avg": D.AGG_AVG, "count_star": D.AGG_COUNT_STAR}
_AGG_MODES = {"Partial": D.AGG_PARTIAL, "Final": D.AGG_FINAL, "FinalPartitioned": D.AGG_FINAL_PARTITIONED, "Single": D.AGG_SINGLE,
              "SinglePartitioned": D.AGG_SINGLE_PARTITIONED, "PartialReduce": D.AGG_PARTIAL_REDUCE}


class AggregateExpr:
    """AggregateFunctionExpr: func(arg) [FILTER (WHERE filter)] AS alias"""

    def __init__(self, func: str, arg: Optional[str], alias: Optional[str] = None, filter: Optional[str] = None):
        self.func, self.arg, self.filter = func.lower(), arg, filter
        self.alias = alias or f"{func.upper()}({arg or '*'})"

    def value_type(self, t: Optional[pa.DataType]) -> pa.DataType:
        if self.func in ("count", "count_star"):
            return pa.int64()
        if self.func == "avg":
            return pa.float64()
        if self.func == "sum":  # Sum::return_type, functions-aggregate/src/sum.rs:232-261
            if pa.types.is_decimal128(t):
                return pa.decimal128(min(38, t.precision + 10), t.scale)
            return pa.float64() if pa.types.is_floating(t) else pa.uint64() if pa.types.is_unsigned_integer(t) else pa.int64()
        return t

    def state_fields(self, t: Optional[pa.DataType]) -> List[pa.Field]:
        if self.func == "avg":  # [count, sum] (aggregates/mod.rs:3591-3700 snapshots)
            return [pa.field(f"{self.alias}[count]", pa.uint64()), pa.field(f"{self.alias}[sum]", pa.float64())]
        return [pa.field(f"{self.alias}[{self.func}]", self.value_type(t))]


class GpuAggregateExec(ExecutionPlan):
    """AggregateExec::try_new(mode, group_by, aggr_expr, filter_expr, input, input_schema) (aggregates/mod.rs:930)."""

    def __init__(self, mode: str, group_by: Sequence[str], aggr_expr: Sequence[AggregateExpr], input: ExecutionPlan,
                 input_schema: Optional[pa.Schema] = None, capacity_hint: int = 0):
        self.mode, self.group_by, self.aggr_expr, self.input = mode, list(group_by), list(aggr_expr), input
        self.input_schema = input_schema or input.schema  # schema of the RAW input (needed in Final modes for value types)
        self.capacity_hint = capacity_hint
        self.state_input = mode in ("Final", "FinalPartitioned", "PartialReduce")
        self.state_output = mode in ("Partial", "PartialReduce")
        gfields = [input.schema.field(g) for g in self.group_by]
        afields: List[pa.Field] = []
        for a in self.aggr_expr:
            t = self.input_schema.field(a.arg).type if a.arg is not None else None
            if self.state_output:
                afields += a.state_fields(t)
            else:
                afields.append(pa.field(a.alias, a.value_type(t)))
        self.schema = pa.schema([pa.field(f.name, f.type, True) for f in gfields] + afields)
        self._metrics = {}

    def children(self): return [self.input]

    def execute(self, ctx):
        isch = self.input.schema
        types = [type_id(f.type) for f in isch]
        gcols = [isch.get_field_index(g) for g in self.group_by]
        aggs = []
        for a in self.aggr_expr:
            if self.state_input:
                aggs.append((_AGG_FUNCS[a.func], -1, -1))
            else:
                aggs.append((_AGG_FUNCS[a.func], isch.get_field_index(a.arg) if a.arg is not None else -1,
                             isch.get_field_index(a.filter) if a.filter else -1))
        if self.state_input:  # layout contract: [group cols..., state cols...] in order
            assert gcols == list(range(len(gcols))), "state input must start with the group columns"
        op = D.AggHandle(ctx.gpu, types, gcols, aggs, _AGG_MODES[self.mode], ctx.config.batch_size, self.capacity_hint)
        try:
            for rb in self.input.execute(ctx):
                op.push_arrow(rb)
            op.finish()
            bs = ctx.config.batch_size
            for rb in _drain(op, self.schema):   # one big batch sliced by batch_size (aggregate_hash_table/common.rs:290)
                for s in range(0, rb.num_rows, bs):
                    yield rb.slice(s, bs)
            self._metrics = {k: op.metric(k) for k in ("num_groups", "input_rows", "output_rows", "rehashes")}
        finally:
            op.close()

    def metrics(self): return self._metrics


# ---------------------------------------------------------------------------------------------
# pipeline fusion — the executable twin of the second optimizer rule of INTEGRATION.md §2a
# ---------------------------------------------------------------------------------------------
class _Scan:
    """the probe-side chain of one pipeline: source plan, predicate over the source schema, probe stages, names visible downstream"""

    def __init__(self, source: ExecutionPlan, predicate: Optional[Expr] = None):
        self.source, self.predicate = source, predicate
        self.stages: List[Tuple[int, str, "GpuPipelineExec"]] = []     # (stage kind, probe key column, build pipeline)
        self.visible: List[str] = [f.name for f in source.schema]       # column names the operators above may still reference

    def virtual_schema(self) -> pa.Schema:
        fields = list(self.source.schema)
        for kind, _, build in self.stages:
            if kind == D.STAGE_INNER:
                fields += [build.scan_field(n) for n in build.payload]
        return pa.schema(fields)


def _as_scan(plan: ExecutionPlan) -> Optional[_Scan]:
    """[ProjectionExec(columns only)]* over [FilterExec]? over [HashJoinExec(RightSemi / RightAnti / Inner, one key, fusable build)]* over a source"""
    if isinstance(plan, GpuProjectionExec):
        if not all(isinstance(e, Column) and e.name == name for e, name in plan.exprs):
            return None
        sc = _as_scan(plan.input)
        if sc is not None:
            sc.visible = [name for _, name in plan.exprs]
        return sc
    if isinstance(plan, GpuFilterExec):
        if plan.fetch is not None:
            return None
        inner = plan.input
        if isinstance(inner, (GpuFilterExec, GpuHashJoinExec, GpuProjectionExec, GpuAggregateExec)):
            return None                                   # predicates are fused over the source columns only
        sc = _Scan(inner, plan.predicate)
        if plan.projection is not None:
            sc.visible = [inner.schema.field(i).name for i in plan.projection]
        return sc
    if isinstance(plan, GpuHashJoinExec):
        if plan.join_type not in ("RightSemi", "RightAnti", "Inner") or len(plan.on) != 1 or plan.filter is not None or plan.null_aware or plan.null_equality != "NullEqualsNothing":
            return None
        sc = _as_scan(plan.right)
        if sc is None or plan.on[0][1] not in [f.name for f in sc.source.schema]:
            return None
        kind = {"RightSemi": D.STAGE_SEMI, "RightAnti": D.STAGE_ANTI, "Inner": D.STAGE_INNER}[plan.join_type]
        payload = [f.name for f in plan.left.schema if f.name != plan.on[0][0]] if kind == D.STAGE_INNER else []
        build = _as_build(plan.left, plan.on[0][0], payload)
        if build is None:
            return None
        sc.stages.append((kind, plan.on[0][1], build))
        names = [f.name for f in plan.schema]
        sc.visible = names
        return sc
    if isinstance(plan, (GpuAggregateExec,)):
        return None
    return _Scan(plan)


def _as_build(plan: ExecutionPlan, key: str, payload: List[str]) -> Optional["GpuPipelineExec"]:
    sc = _as_scan(plan)
    if sc is None or len(sc.stages) >= 3:
        return None
    vs = sc.virtual_schema()
    if vs.get_field_index(key) < 0 or vs.get_field_index(key) >= len(sc.source.schema) or any(vs.get_field_index(n) < 0 for n in payload):
        return None
    bits = sum(D.WIDTH[type_id(vs.field(n).type)] * 8 for n in payload)
    if bits > 64:
        return None
    return GpuPipelineExec(sc, sink="build", key=key, payload=payload)


class GpuPipelineExec(ExecutionPlan):
    """One fused pipeline (dfgpu_pipeline): source -> predicate -> probe stages -> {build | aggregate | output}.  Built by fuse_pipelines()."""

    def __init__(self, scan: _Scan, sink: str, key: Optional[str] = None, payload: Sequence[str] = (), group_by: Sequence[str] = (),
                 aggs: Sequence[Tuple[str, Optional[Expr], str]] = (), mode: str = "Single", out_schema: Optional[pa.Schema] = None):
        self.scan, self.sink, self.key, self.payload, self.group_by, self.aggs, self.mode = scan, sink, key, list(payload), list(group_by), list(aggs), mode
        self.schema = out_schema if out_schema is not None else pa.schema([])
        self.n_acc_words = 0
        self._metrics = {}

    def children(self): return [self.scan.source] + [b for _, _, b in self.scan.stages]
    def scan_field(self, name: str) -> pa.Field: return self.scan.virtual_schema().field(name)
    def metrics(self): return self._metrics

    # ---- build side: run the pipeline into a dfgpu_lookup ----
    def build_lookup(self, ctx: TaskContext) -> D.Lookup:
        vs = self.scan.virtual_schema()
        batches = list(self.scan.source.execute(ctx))
        key_range = None
        ktype = vs.field(self.key).type
        if not self.payload and self.n_acc_words == 0 and pa.types.is_integer(ktype) and batches:   # statistics: the bounds collect_left_input tracks
            import pyarrow.compute as pc
            mm = [pc.min_max(b.column(self.scan.source.schema.get_field_index(self.key))) for b in batches if b.num_rows]
            lo = [m["min"].as_py() for m in mm if m["min"].is_valid]; hi = [m["max"].as_py() for m in mm if m["max"].is_valid]
            if lo:
                key_range = (min(lo), max(hi))
        look = D.Lookup(ctx.gpu, type_id(ktype), [type_id(vs.field(n).type) for n in self.payload], key_range=key_range, n_acc_words=self.n_acc_words)
        pipe, keep = self._make_pipeline(ctx)
        try:
            pipe.sink_build(look, vs.get_field_index(self.key), [vs.get_field_index(n) for n in self.payload])
            for rb in batches:
                pipe.push_arrow(rb)
            pipe.finish()
            self._metrics = {"input_rows": pipe.metric("input_rows"), "build_rows": pipe.metric("sink_rows"), "lookup_mode": look.metric("mode")}
        finally:
            pipe.close()
            for l in keep:
                l.close()
        return look

    def _make_pipeline(self, ctx: TaskContext):
        ssch = self.scan.source.schema
        nodes = None
        if self.scan.predicate is not None:
            nodes = []
            self.scan.predicate.rpn(ssch, nodes)
        stages, keep = [], []
        for kind, pkey, build in self.scan.stages:
            look = build.build_lookup(ctx)                      # the pipeline breaker: WaitBuildSide (hash_join/stream.rs:117-140)
            keep.append(look)
            stages.append((kind, ssch.get_field_index(pkey), look))
        return D.Pipeline(ctx.gpu, [type_id(f.type) for f in ssch], nodes, stages), keep

    def execute(self, ctx):
        assert self.sink == "aggregate", "build pipelines are driven by their consumer"
        vs = self.scan.virtual_schema()
        pipe, keep = self._make_pipeline(ctx)
        try:
            aggs = []
            for func, expr, _ in self.aggs:
                if expr is None:
                    aggs.append((_AGG_FUNCS[func], None))
                else:
                    nodes: list = []
                    expr.rpn(vs, nodes)
                    aggs.append((_AGG_FUNCS[func], nodes))
            pipe.sink_aggregate([vs.get_field_index(g) for g in self.group_by], aggs, _AGG_MODES[self.mode], 0)
            for rb in self.scan.source.execute(ctx):
                pipe.push_arrow(rb)
            pipe.finish()
            bs = ctx.config.batch_size
            for rb in _drain(pipe, self.schema):
                for s in range(0, rb.num_rows, bs):
                    yield rb.slice(s, bs)
            self._metrics = {k: pipe.metric(k) for k in ("input_rows", "sink_rows", "num_groups")}
        finally:
            pipe.close()
            for l in keep:
                l.close()


def fuse_pipelines(plan: ExecutionPlan) -> ExecutionPlan:
    """PhysicalOptimizerRule twin: AggregateExec(Single / SinglePartitioned / Partial) over [ProjectionExec] over
    HashJoinExec(Inner) whose GROUP BY is the probe key plus build-side columns becomes ONE GpuPipelineExec; its build side (filters, semi
    joins, column projections) becomes build pipelines.  Anything else is returned unchanged (the unfused Gpu*Exec operators run)."""
    if not isinstance(plan, GpuAggregateExec) or plan.mode not in ("Single", "SinglePartitioned", "Partial") or not plan.group_by:
        return plan
    below, proj = plan.input, None
    if isinstance(below, GpuProjectionExec):
        proj, below = below, below.input
    if not isinstance(below, GpuHashJoinExec) or below.join_type != "Inner":
        return plan
    sc = _as_scan(below)
    if sc is None or not sc.stages or sc.stages[-1][0] != D.STAGE_INNER:
        return plan
    kind, pkey, build = sc.stages[-1]
    vs = sc.virtual_schema()
    exprs = {name: e for e, name in proj.exprs} if proj is not None else {f.name: Column(f.name) for f in below.schema}
    # group keys: the probe key (or the equal build key) and payload fields of the LAST inner stage
    group = []
    for g in plan.group_by:
        e = exprs.get(g)
        if not isinstance(e, Column):
            return plan
        n = pkey if e.name == build.key else e.name
        if n != pkey and n not in build.payload:
            return plan
        group.append(n)
    if pkey not in group:
        return plan
    aggs = []
    for a in plan.aggr_expr:
        if a.filter is not None:
            return plan
        e = None if a.arg is None else exprs.get(a.arg)
        if a.arg is not None and e is None:
            return plan
        if a.func == "avg" and e is not None and e.data_type(vs) != pa.float64():
            return plan
        aggs.append((a.func, e, a.alias))
    try:
        for _, e, _ in aggs:
            if e is not None:
                e.rpn(vs, [])                              # every referenced name must exist in the virtual schema
    except KeyError:
        return plan
    build.n_acc_words = 1 + sum(2 if f == "avg" else 1 for f, _, _ in aggs) + sum(1 for f, _, _ in aggs if f in ("sum", "min", "max"))
    if build.n_acc_words > 12:
        return plan
    return GpuPipelineExec(sc, sink="aggregate", group_by=group, aggs=aggs, mode=plan.mode, out_schema=plan.schema)


def collect(plan: ExecutionPlan, ctx: Optional[TaskContext] = None) -> List[pa.RecordBatch]:
    """physical-plan/src/execution_plan.rs:1752"""
    ctx = ctx or TaskContext()
    return [b for b in plan.execute(ctx) if b.num_rows > 0]
