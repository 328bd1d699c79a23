"""The fusion rule as a PLANNING decision (no GPU needed: nothing is executed).  `fuse_pipelines` is the executable twin of the
PhysicalOptimizerRule a maintainer registers (reference hook: datafusion/session/src/physical_optimizer.rs:52): it
must collapse the reference's Q3 operator tree (sqllogictest/test_files/tpch/plans/q3.slt.part:60-76) into one GpuPipelineExec over
build pipelines, and hand back — untouched — every plan whose semantics the fused kernel cannot carry (the CPU / unfused operators run)."""
import datetime

import numpy as np
import pyarrow as pa

from datafusion_b200 import capi as D
from datafusion_b200.exec import (AggregateExpr, GpuAggregateExec, GpuFilterExec, GpuHashJoinExec, GpuPipelineExec, GpuProjectionExec, MemoryExec, col,
                                  fuse_pipelines, lit)

CUT = datetime.date(1995, 3, 15)


def tables():
    customer = pa.table({"c_custkey": np.arange(1, 9, dtype=np.int64), "c_mktsegment": np.array([1, 0, 1, 2, 1, 3, 4, 1], np.int32)})
    orders = pa.table({"o_orderkey": np.arange(1, 17, dtype=np.int64), "o_custkey": (np.arange(16, dtype=np.int64) % 8) + 1,
                       "o_orderdate": pa.array(np.arange(9000, 9016, dtype=np.int32)).cast(pa.date32()), "o_shippriority": np.zeros(16, np.int32)})
    lineitem = pa.table({"l_orderkey": (np.arange(40, dtype=np.int64) % 16) + 1, "l_extendedprice": np.arange(40, dtype=np.int64) * 100 + 90_000,
                         "l_discount": np.arange(40, dtype=np.int64) % 11, "l_shipdate": pa.array(np.arange(9100, 9140, dtype=np.int32)).cast(pa.date32())})
    return customer, orders, lineitem


def q3_parts(join_type="Inner", semi_type="RightSemi"):
    customer, orders, lineitem = tables()
    mem = lambda t: MemoryExec(t.to_batches(), t.schema)
    c = GpuFilterExec(col("c_mktsegment") == lit(1, pa.int32()), mem(customer), projection=[0])
    o = GpuFilterExec(col("o_orderdate") < lit(CUT, pa.date32()), mem(orders))
    semi = GpuHashJoinExec(c, o, [("c_custkey", "o_custkey")], semi_type)
    semi_p = GpuProjectionExec([(col("o_orderkey"), "o_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority")], semi)
    l = GpuFilterExec(col("l_shipdate") > lit(CUT, pa.date32()), mem(lineitem), projection=[0, 1, 2])
    inner = GpuHashJoinExec(semi_p, l, [("o_orderkey", "l_orderkey")], join_type, projection=[1, 2, 3, 4, 5] if join_type == "Inner" else None)
    return semi_p, l, inner


def revenue_projection(inner):
    return GpuProjectionExec([(col("l_orderkey"), "l_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority"),
                              (col("l_extendedprice") * (lit(100, pa.int64()) - col("l_discount")), "rev")], inner)


def test_rule_collapses_the_q3_tree_into_one_pipeline_over_two_build_pipelines():
    _, _, inner = q3_parts()
    rev = revenue_projection(inner)
    for mode in ("SinglePartitioned", "Single", "Partial"):
        agg = GpuAggregateExec(mode, ["l_orderkey", "o_orderdate", "o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev)
        fused = fuse_pipelines(agg)
        assert isinstance(fused, GpuPipelineExec) and fused.schema == agg.schema
        stages = fused.scan.stages
        assert len(stages) == 1 and stages[0][0] == D.STAGE_INNER and stages[0][1] == "l_orderkey"          # lineitem scan -> Inner probe on l_orderkey
        orders_build = stages[0][2]
        assert orders_build.key == "o_orderkey" and orders_build.payload == ["o_orderdate", "o_shippriority"]
        assert len(orders_build.scan.stages) == 1 and orders_build.scan.stages[0][0] == D.STAGE_SEMI        # orders scan -> RightSemi probe of the customer key set
        assert orders_build.scan.stages[0][1] == "o_custkey" and orders_build.scan.stages[0][2].key == "c_custkey"
        assert orders_build.n_acc_words == 1 + 1 + 1                                                          # row counter, SUM, SUM's non-null counter


def test_rule_accepts_more_aggregates_and_count_star():
    _, _, inner = q3_parts()
    rev = revenue_projection(inner)
    agg = GpuAggregateExec("Single", ["l_orderkey", "o_orderdate"], [AggregateExpr("sum", "rev", "revenue"), AggregateExpr("count", None, "n"), AggregateExpr("max", "rev", "top")], rev)
    fused = fuse_pipelines(agg)
    assert isinstance(fused, GpuPipelineExec) and [a[0] for a in fused.aggs] == ["sum", "count", "max"]


def test_rule_leaves_alone_what_the_fused_kernel_cannot_carry():
    _, _, inner = q3_parts()
    rev = revenue_projection(inner)
    same = lambda p: fuse_pipelines(p) is p
    # a GROUP BY the probe key does not determine (the group id would not be the build row)
    assert same(GpuAggregateExec("Single", ["o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev))
    # a computed group key
    shifted = GpuProjectionExec([(col("l_orderkey") + lit(1, pa.int64()), "k"), (col("l_extendedprice"), "rev")], inner)
    assert same(GpuAggregateExec("Single", ["k"], [AggregateExpr("sum", "rev", "revenue")], shifted))
    # Final / FinalPartitioned merge states: nothing to fuse with a probe
    part = GpuAggregateExec("Partial", ["l_orderkey", "o_orderdate", "o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev)
    assert same(GpuAggregateExec("Final", ["l_orderkey", "o_orderdate", "o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], part, input_schema=rev.schema))
    # an aggregate FILTER clause
    flt = GpuProjectionExec([(col("l_orderkey"), "l_orderkey"), (col("l_extendedprice"), "rev"), (col("l_discount") > lit(3, pa.int64()), "keep")], inner)
    assert same(GpuAggregateExec("Single", ["l_orderkey"], [AggregateExpr("sum", "rev", "revenue", filter="keep")], flt))
    # not an aggregate on top; no GROUP BY
    assert same(rev) and same(inner)
    assert same(GpuAggregateExec("Single", [], [AggregateExpr("sum", "rev", "revenue")], rev))
    # the topmost join is not Inner: unmatched probe rows must come out NULL-padded
    _, _, right = q3_parts(join_type="Right")
    r2 = GpuProjectionExec([(col("l_orderkey"), "l_orderkey"), (col("l_extendedprice"), "rev")], right)
    assert same(GpuAggregateExec("Single", ["l_orderkey"], [AggregateExpr("sum", "rev", "revenue")], r2))
