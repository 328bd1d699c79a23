"""The reference-facing operator mirror (datafusion_b200/exec.py) driven like the reference's own tests:
TestMemoryExec-style inputs, `collect(plan)`, snapshot comparison — every batch crossing the C ABI as an
Arrow C Data Interface struct array (push_arrow / export_arrow)."""
import numpy as np
import pyarrow as pa
import pytest

from datafusion_b200.exec import (AggregateExpr, DictionaryDecodeExec, DictionaryEncodeExec, GpuAggregateExec, GpuFilterExec, GpuHashJoinExec, JoinFilter, MemoryExec, SessionConfig, TaskContext, plan_string_dictionary,
                                  col, collect, lit)
from harness import load_golden

pytestmark = pytest.mark.gpu
KAT = load_golden("hash_join_kat.json")["cases"]
MISC = load_golden("misc_kat.json")


def build_table(cols):
    """build_table / build_table_two_cols of hash_join/exec.rs:2965-2999: Int32 columns"""
    return pa.RecordBatch.from_arrays([pa.array(v, type=pa.int32()) for _, v in cols], names=[n for n, _ in cols])


def table_rows(batches):
    t = pa.Table.from_batches(batches) if batches else None
    return [] if t is None else [list(r.values()) for r in t.to_pylist()]


def srt(rows):
    return sorted(rows, key=lambda r: [(x is None, x if x is not None else 0) for x in r])


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_join_snapshots_through_arrow_boundary(gpu_ctx, case):
    left, right = build_table(case["left"]), build_table(case["right"])
    # duplicate column names across sides (e.g. b1/b1) are legal in the reference; make them unique for pyarrow
    rnames = [n if n not in left.schema.names else n + "_r" for n in right.schema.names]
    on = [(l, rnames[right.schema.names.index(r)]) for l, r in case["on"]]
    right = pa.RecordBatch.from_arrays(right.columns, names=rnames)
    for batch_size in (8192, 2):
        for phj in (True, False):
            cfg = SessionConfig(batch_size=batch_size, perfect_hash_join_small_build_threshold=819200 if phj else 0,
                                perfect_hash_join_min_key_density=0.0 if phj else float("inf"))
            jf = None
            if case.get("filter"):   # prepare_join_filter (exec.rs:5556-5583): JoinFilter(c@0 > c@1, [Left:2, Right:2])
                jf = JoinFilter(col("f0") > col("f1"), [("left", 2), ("right", 2)])
            join = GpuHashJoinExec(MemoryExec([left] * case.get("left_repeat", 1)), MemoryExec([right] * case.get("right_repeat", 1)), on,
                                   case["join_type"], case["null_equality"], filter=jf)
            got = table_rows(collect(join, TaskContext(cfg, gpu_ctx)))
            assert len(join.schema) == len(case["header"])
            exp = case["expected"]
            if case["sorted"] or case["join_type"] in ("Left", "Right", "Full", "LeftSemi", "LeftAnti", "LeftMark"):
                assert srt(got) == srt(exp), case["name"]
            else:
                assert got == exp, case["name"]   # batches_to_string: exact order (exec.rs:3339 "preserve both inputs order")
            if len(on) == 1:
                assert join.metrics()["array_map_created_count"] == (1 if phj else 0)


def test_filter_exec_like_reference(gpu_ctx, task_ctx):
    rng = np.random.default_rng(0)
    n = 100_000
    batches = [pa.RecordBatch.from_arrays([pa.array(rng.integers(0, 1000, 8192), type=pa.int64()),
                                           pa.array(rng.integers(0, 100, 8192).astype(np.int32), mask=rng.random(8192) < 0.1)], names=["x", "y"])
               for _ in range(n // 8192)]
    plan = GpuFilterExec((col("x") > 500) & ((col("y") < 50) | col("y").is_null()), MemoryExec(batches))
    got = pa.Table.from_batches(collect(plan, task_ctx))
    t = pa.Table.from_batches(batches)
    import pyarrow.compute as pc
    mask = pc.and_kleene(pc.greater(t["x"], 500), pc.or_kleene(pc.less(t["y"], 50), pc.is_null(t["y"])))
    exp = t.filter(mask)
    assert got.equals(exp)
    assert all(b.num_rows == 8192 for b in collect(plan, task_ctx)[:-1])
    with pytest.raises(ValueError, match="must return BOOLEAN"):
        GpuFilterExec(col("x") + 1, MemoryExec(batches))
    # date32 < literal (TPC-H Q3: o_orderdate < 1995-03-15)
    import datetime
    d = pa.array(rng.integers(8000, 10500, 5000).astype(np.int32), type=pa.int32()).cast(pa.date32())
    plan = GpuFilterExec(col("d") < lit(datetime.date(1995, 3, 15), pa.date32()), MemoryExec([pa.RecordBatch.from_arrays([d], names=["d"])]))
    got = pa.Table.from_batches(collect(plan, task_ctx))
    assert got["d"].to_pylist() == [x for x in d.to_pylist() if x < datetime.date(1995, 3, 15)]


def test_aggregate_exec_partial_final_like_check_aggregates(gpu_ctx, task_ctx):
    m = MISC["aggregate_some_data"]
    schema = pa.schema([("a", pa.uint32()), ("b", pa.float64())])
    batches = [pa.RecordBatch.from_arrays([pa.array(b["a"], type=pa.uint32()), pa.array(b["b"], type=pa.float64())], schema=schema) for b in m["batches"]]
    aggs = [AggregateExpr("avg", "b", "AVG(b)")]
    partial = GpuAggregateExec("Partial", ["a"], aggs, MemoryExec(batches))
    assert partial.schema.names == ["a", "AVG(b)[count]", "AVG(b)[sum]"]          # aggregates/mod.rs:3636-3646 snapshot header
    pres = collect(partial, task_ctx)
    got = srt(table_rows(pres))
    assert got == [[a, c, s] for a, c, s in zip(m["partial"]["a"], m["partial"]["count"], m["partial"]["sum"])]
    final = GpuAggregateExec("Final", ["a"], aggs, MemoryExec(pres + pres, partial.schema), input_schema=schema)
    fres = srt(table_rows(collect(final, task_ctx)))
    assert fres == [[a, v] for a, v in zip(m["final_avg"]["a"], m["final_avg"]["avg"])]
    assert final.metrics()["output_rows"] == 3


def test_poll_ready_is_non_blocking_and_becomes_true(gpu_ctx):
    """dfgpu_poll_ready: the waker side of the async contract (execution_plan.rs:549-563) — never blocks, 1 once the stream has drained"""
    import time
    from datafusion_b200 import capi as D
    gpu_ctx.sync()
    assert gpu_ctx.poll_ready() is True
    buf = gpu_ctx.generate_i64(D.GEN_SPLITMIX, 1, 0, 0, 0, 200_000_000)      # queued asynchronously on the ctx stream
    t0 = time.perf_counter()
    first = gpu_ctx.poll_ready()
    assert time.perf_counter() - t0 < 0.05                                     # answered without waiting for the kernel
    while not gpu_ctx.poll_ready():
        time.sleep(0.0005)
    assert gpu_ctx.poll_ready() is True and first in (True, False)
    buf.free()


def test_decimal_money_through_the_exec_twin(gpu_ctx, task_ctx):
    """Decimal128(15,2) columns through the Arrow boundary: FilterExec on a decimal predicate, the revenue expression's type, SUM's
    Decimal128(25,2) result — against pyarrow on the same record batches"""
    import decimal
    import pyarrow.compute as pc
    rng = np.random.default_rng(3)
    n = 50_000
    price = pa.array([decimal.Decimal(int(v)).scaleb(-2) for v in rng.integers(90_000, 10_500_000, n)], pa.decimal128(15, 2))
    disc = pa.array([None if rng.random() < 0.05 else decimal.Decimal(int(v)).scaleb(-2) for v in rng.integers(0, 11, n)], pa.decimal128(15, 2))
    flag = pa.array(rng.integers(0, 7, n), pa.int32())
    rb = pa.record_batch([price, disc, flag], names=["p", "d", "k"])
    src = MemoryExec([rb.slice(0, 20_000), rb.slice(20_000)])
    f = GpuFilterExec(col("d") >= lit(decimal.Decimal("0.05"), pa.decimal128(15, 2)), src)
    assert (col("p") * (lit(1) - col("d"))).data_type(rb.schema) == pa.decimal128(38, 4)
    agg = GpuAggregateExec("Single", ["k"], [AggregateExpr("sum", "p"), AggregateExpr("count", "d")], f)
    assert agg.schema.field(1).type == pa.decimal128(25, 2)
    got = pa.Table.from_batches(list(agg.execute(task_ctx)), schema=agg.schema).sort_by("k")
    keep = pc.fill_null(pc.greater_equal(disc, pa.scalar(decimal.Decimal("0.05"), pa.decimal128(15, 2))), False)
    ref = pa.table({"k": flag, "p": price, "d": disc}).filter(keep).group_by("k").aggregate([("p", "sum"), ("d", "count")]).sort_by("k")
    assert got.column(0).to_pylist() == ref["k"].to_pylist()
    assert got.column(1).to_pylist() == ref["p_sum"].to_pylist() and got.column(2).to_pylist() == ref["d_count"].to_pylist()


def test_exported_arrow_children_outlive_the_parent(gpu_ctx):
    """Arrow C Data Interface: a child moved out of the exported struct array stays valid after the parent (and the dfgpu batch) are
    released — pyarrow's import does exactly that per column"""
    import gc
    from datafusion_b200 import capi as D
    x = np.arange(100_000, dtype=np.int64)
    f = D.FilterHandle(gpu_ctx, [D.INT64], [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, 49_999, 0.0), (D.EXPR_BINARY, D.OP_GT, 0, 0, 0, 0.0)], None, 1 << 20, -1)
    f.push_host([D.HostColumn(x)]); f.finish()
    outs = f.drain(host=True)
    rbs = [b.to_arrow() for b in outs]
    cols = [rb.column(0) for rb in rbs]          # the children alone
    for b in outs:
        b.release()
    del outs, rbs
    f.close()
    gc.collect()
    junk = [np.ones(1 << 20, np.int64) for _ in range(8)]     # churn the host allocator
    assert pa.concat_arrays(cols).to_numpy().tolist() == list(range(50_000, 100_000))
    del junk


def test_string_keys_through_the_exec_twin(gpu_ctx, task_ctx):
    """Utf8 and Dictionary(Int32, Utf8) key columns: DictionaryEncodeExec -> GpuHashJoinExec -> GpuAggregateExec -> DictionaryDecodeExec,
    one plan-wide code space; against pyarrow joining and grouping on the strings themselves"""
    rng = np.random.default_rng(12)
    words = ["AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD", "", "büilding", "x" * 200]
    nl, nr = 400, 9000
    lseg = [None if rng.random() < 0.05 else words[int(rng.integers(0, 6))] for _ in range(nl)]            # plain Utf8, 6 of the 8 words
    left = pa.record_batch([pa.array(lseg, pa.string()), pa.array(np.arange(nl), pa.int64())], names=["seg", "lid"])
    rbatches, rseg_all, rv_all = [], [], []
    for part in range(3):                                                                                   # dictionary arrays, another dictionary per batch
        local = [words[i] for i in rng.permutation(len(words))[: int(rng.integers(2, 9))]]
        idx = rng.integers(0, len(local), nr // 3)
        mask = rng.random(nr // 3) < 0.04
        seg = pa.DictionaryArray.from_arrays(pa.array(idx.astype(np.int32), mask=mask), pa.array(local, pa.string()))
        v = rng.integers(-100, 100, nr // 3).astype(np.int64)
        rbatches.append(pa.record_batch([seg, pa.array(v)], names=["seg_r", "v"]))
        rseg_all += seg.to_pylist(); rv_all += v.tolist()
    dictionary_of = plan_string_dictionary()
    join = GpuHashJoinExec(DictionaryEncodeExec(MemoryExec([left]), dictionary_of), DictionaryEncodeExec(MemoryExec(rbatches), dictionary_of), [("seg", "seg_r")], "Inner")
    agg = GpuAggregateExec("Single", ["seg"], [AggregateExpr("sum", "v"), AggregateExpr("count_star", None), AggregateExpr("sum", "lid")], join)
    out = DictionaryDecodeExec(agg, ["seg"], dictionary_of)
    assert out.schema.field(0).type == pa.string()
    got = pa.Table.from_batches(list(out.execute(task_ctx)), schema=out.schema).sort_by("seg")
    lt = pa.table({"seg": pa.array(lseg, pa.string()), "lid": np.arange(nl)})
    rt = pa.table({"seg_r": pa.array(rseg_all, pa.string()), "v": rv_all})
    ref = lt.join(rt, keys="seg", right_keys="seg_r", join_type="inner").group_by("seg").aggregate([("v", "sum"), ([], "count_all"), ("lid", "sum")]).sort_by("seg")
    assert got.column(0).to_pylist() == ref["seg"].to_pylist() and len(ref) >= 5
    assert got.column(1).to_pylist() == ref["v_sum"].to_pylist()
    assert got.column(2).to_pylist() == ref["count_all"].to_pylist() and got.column(3).to_pylist() == ref["lid_sum"].to_pylist()
    dictionary_of(task_ctx).close()
