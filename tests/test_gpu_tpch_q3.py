"""BASELINE config C4 as a parity case: the TPC-H Q3-shaped pipeline of the reference's physical plan
(sqllogictest/test_files/tpch/plans/q3.slt.part:60-76)
    FilterExec(c_mktsegment = BUILDING) -> HashJoinExec RightSemi (c_custkey = o_custkey) over FilterExec(o_orderdate < 1995-03-15)
    -> HashJoinExec Inner (o_orderkey = l_orderkey) over FilterExec(l_shipdate > 1995-03-15)
    -> AggregateExec gby [l_orderkey, o_orderdate, o_shippriority] SUM(l_extendedprice * (100 - l_discount))
on synthetic TPC-H-shaped tables (int64 fixed-point money), executed operator by operator on the
GPU through the Arrow boundary and compared with an independent numpy/pyarrow evaluation."""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from datafusion_b200.exec import (AggregateExpr, GpuAggregateExec, GpuFilterExec, GpuHashJoinExec, GpuPipelineExec, GpuProjectionExec, MemoryExec, SessionConfig,
                                  TaskContext, col, collect, fuse_pipelines, lit)

pytestmark = pytest.mark.gpu
EPOCH = datetime.date(1970, 1, 1)
D0, D1 = (datetime.date(1992, 1, 1) - EPOCH).days, (datetime.date(1998, 8, 2) - EPOCH).days
CUT = datetime.date(1995, 3, 15)


def gen_tables(sf, seed=0):
    rng = np.random.default_rng(seed)
    nc, no = int(150_000 * sf), int(1_500_000 * sf)
    c_custkey = np.arange(1, nc + 1, dtype=np.int64)
    c_mktsegment = rng.integers(0, 5, nc).astype(np.int32)                       # code 1 = 'BUILDING' (Utf8View in the real schema)
    o_orderkey = (np.arange(no, dtype=np.int64) // 8) * 32 + (np.arange(no, dtype=np.int64) % 8) + 1   # sparse keys: 8 of every 32 (dbgen)
    o_custkey = rng.integers(1, nc * 2 // 3 + 1, no).astype(np.int64)            # a third of the customers have no orders
    o_orderdate = rng.integers(D0, D1 + 1, no).astype(np.int32)
    o_shippriority = np.zeros(no, np.int32)
    nl = rng.integers(1, 8, no)
    l_orderkey = np.repeat(o_orderkey, nl)
    n = len(l_orderkey)
    l_extendedprice = rng.integers(90_000, 10_500_000, n).astype(np.int64)       # cents
    l_discount = rng.integers(0, 11, n).astype(np.int64)                         # hundredths
    l_shipdate = (np.repeat(o_orderdate, nl) + rng.integers(1, 122, n)).astype(np.int32)
    perm = rng.permutation(n)
    customer = pa.table({"c_custkey": c_custkey, "c_mktsegment": c_mktsegment})
    orders = pa.table({"o_orderkey": o_orderkey, "o_custkey": o_custkey, "o_orderdate": pa.array(o_orderdate).cast(pa.date32()), "o_shippriority": o_shippriority})
    lineitem = pa.table({"l_orderkey": l_orderkey[perm], "l_extendedprice": l_extendedprice[perm], "l_discount": l_discount[perm],
                         "l_shipdate": pa.array(l_shipdate[perm]).cast(pa.date32())})
    return customer, orders, lineitem


def q3_expected(customer, orders, lineitem):
    cut = (CUT - EPOCH).days
    c = customer.to_pandas(); o = orders.to_pandas(); l = lineitem.to_pandas()
    o["o_orderdate"] = np.asarray(orders["o_orderdate"].cast(pa.int32())); l["l_shipdate"] = np.asarray(lineitem["l_shipdate"].cast(pa.int32()))
    ck = set(c.loc[c.c_mktsegment == 1, "c_custkey"])
    o = o[(o.o_orderdate < cut) & o.o_custkey.isin(ck)]
    l = l[l.l_shipdate > cut]
    j = l.merge(o, left_on="l_orderkey", right_on="o_orderkey")
    j["rev"] = j.l_extendedprice * (100 - j.l_discount)
    g = j.groupby(["l_orderkey", "o_orderdate", "o_shippriority"], as_index=False)["rev"].sum()
    return sorted(zip(g.l_orderkey.tolist(), g.o_orderdate.tolist(), g.o_shippriority.tolist(), g.rev.tolist()))


@pytest.mark.parametrize("sf,batch_rows", [(0.002, 1000), (0.05, 8192)])
def test_q3_pipeline_matches_independent_evaluation(gpu_ctx, sf, batch_rows):
    customer, orders, lineitem = gen_tables(sf)
    ctx = TaskContext(SessionConfig(), gpu_ctx)
    mem = lambda t: MemoryExec(t.to_batches(max_chunksize=batch_rows), t.schema)
    c = GpuFilterExec(col("c_mktsegment") == lit(1, pa.int32()), mem(customer), projection=[0])
    o = GpuFilterExec(col("o_orderdate") < lit(CUT, pa.date32()), mem(orders))
    semi = GpuHashJoinExec(c, o, [("c_custkey", "o_custkey")], "RightSemi")           # q3.slt.part:66 (RightSemi)
    semi_p = GpuProjectionExec([(col("o_orderkey"), "o_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority")], semi)
    l = GpuFilterExec(col("l_shipdate") > lit(CUT, pa.date32()), mem(lineitem), projection=[0, 1, 2])
    inner = GpuHashJoinExec(semi_p, l, [("o_orderkey", "l_orderkey")], "Inner", projection=[1, 2, 3, 4, 5])   # q3.slt.part:64
    rev = GpuProjectionExec([(col("l_orderkey"), "l_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority"),
                             (col("l_extendedprice") * (lit(100, pa.int64()) - col("l_discount")), "rev")], inner)
    agg = GpuAggregateExec("SinglePartitioned", ["l_orderkey", "o_orderdate", "o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev)
    out = pa.Table.from_batches(collect(agg, ctx))
    got = sorted(zip(out["l_orderkey"].to_pylist(), np.asarray(out["o_orderdate"].cast(pa.int32())).tolist(), out["o_shippriority"].to_pylist(),
                     out["revenue"].to_pylist()))
    exp = q3_expected(customer, orders, lineitem)
    assert len(exp) > 0 and got == exp
    assert inner.metrics()["array_map_created_count"] == 0 or sf < 0.01   # o_orderkey is sparse (8 of 32): hash path at scale


def q3_plan(customer, orders, lineitem, batch_rows):
    mem = lambda t: MemoryExec(t.to_batches(max_chunksize=batch_rows), t.schema)
    c = GpuFilterExec(col("c_mktsegment") == lit(1, pa.int32()), mem(customer), projection=[0])
    o = GpuFilterExec(col("o_orderdate") < lit(CUT, pa.date32()), mem(orders))
    semi = GpuHashJoinExec(c, o, [("c_custkey", "o_custkey")], "RightSemi")
    semi_p = GpuProjectionExec([(col("o_orderkey"), "o_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority")], semi)
    l = GpuFilterExec(col("l_shipdate") > lit(CUT, pa.date32()), mem(lineitem), projection=[0, 1, 2])
    inner = GpuHashJoinExec(semi_p, l, [("o_orderkey", "l_orderkey")], "Inner", projection=[1, 2, 3, 4, 5])
    rev = GpuProjectionExec([(col("l_orderkey"), "l_orderkey"), (col("o_orderdate"), "o_orderdate"), (col("o_shippriority"), "o_shippriority"),
                             (col("l_extendedprice") * (lit(100, pa.int64()) - col("l_discount")), "rev")], inner)
    return rev, GpuAggregateExec("SinglePartitioned", ["l_orderkey", "o_orderdate", "o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev)


@pytest.mark.parametrize("sf,batch_rows", [(0.002, 1000), (0.05, 8192)])
def test_fusion_rule_collapses_the_q3_plan_into_pipelines(gpu_ctx, sf, batch_rows):
    """the executable twin of the fusion rule: the reference-shaped operator tree of Q3 becomes ONE GpuPipelineExec over two
    build pipelines, fed and drained through the Arrow C Data Interface, and gives the unfused plan's rows"""
    customer, orders, lineitem = gen_tables(sf)
    ctx = TaskContext(SessionConfig(), gpu_ctx)
    rev, agg = q3_plan(customer, orders, lineitem, batch_rows)
    fused = fuse_pipelines(agg)
    assert isinstance(fused, GpuPipelineExec) and fused.schema == agg.schema
    builds = [b for _, _, b in fused.scan.stages]
    assert len(builds) == 1 and builds[0].key == "o_orderkey" and builds[0].payload == ["o_orderdate", "o_shippriority"] and len(builds[0].scan.stages) == 1
    out = pa.Table.from_batches(collect(fused, ctx))
    got = sorted(zip(out["l_orderkey"].to_pylist(), np.asarray(out["o_orderdate"].cast(pa.int32())).tolist(), out["o_shippriority"].to_pylist(), out["revenue"].to_pylist()))
    exp = q3_expected(customer, orders, lineitem)
    assert len(exp) > 0 and got == exp
    assert fused.metrics()["num_groups"] == len(exp) and builds[0].scan.stages[0][2].metrics()["lookup_mode"] == 1     # the customer key set became a bitmap
    # shapes the rule must leave alone: a GROUP BY that the join key does not determine, a Final aggregate
    other = GpuAggregateExec("Single", ["o_shippriority"], [AggregateExpr("sum", "rev", "revenue")], rev)
    assert fuse_pipelines(other) is other
    assert fuse_pipelines(rev) is rev
