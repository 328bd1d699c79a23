"""The CPU arm of the headline benchmark (oracle_bench_q3: the reference's Q3 physical plan, tpch/plans/q3.slt.part:60-76, restated
with T partitions) must give the rows of the oracle's own operator chain (filter_batch -> hash_join RightSemi -> hash_join Inner ->
group_by, each pinned by the reference's vectors) and of an independent pandas evaluation — for every thread count."""
import numpy as np
import pytest

from oracle import oracle as O


def fingerprint(cols):
    m = 2**64
    return [len(cols[0])] + [int(np.asarray(c).astype(np.int64).view(np.uint64).sum(dtype=np.uint64)) % m for c in cols]


def chain(t, cut=O.Q3_CUT):
    c = [(t["c_custkey"], None), (t["c_mktsegment"], None)]
    o = [(t["o_orderkey"], None), (t["o_custkey"], None), (t["o_orderdate"], None), (t["o_shippriority"], None)]
    l = [(t["l_orderkey"], None), (t["l_extendedprice"], None), (t["l_discount"], None), (t["l_shipdate"], None)]
    fc = O.filter_batch(c, (t["c_mktsegment"] == 1, None), [0])
    fo = O.filter_batch(o, (t["o_orderdate"] < cut, None))
    fl = O.filter_batch(l, (t["l_shipdate"] > cut, None), [0, 1, 2])
    so = O.hash_join(fc, fo, [0], [1], [1, 1, 1], [0, 2, 3], join_type=O.J_RIGHT_SEMI)
    j = O.hash_join(so, fl, [0], [0], [0, 0, 1, 1, 1], [1, 2, 0, 1, 2])
    rev = (j[3][0].astype(np.uint64) * (100 - j[4][0]).astype(np.uint64)).view(np.int64)
    keys, res = O.group_by([j[2], j[0], j[1]], [(O.A_SUM, (rev, None), None)])
    return [keys[0][0], keys[1][0], keys[2][0], res[0]["i"]], len(so[0][0]), len(j[0][0])


@pytest.mark.parametrize("sf,threads", [(0.02, 1), (0.02, 3), (0.1, 8)])
def test_bench_q3_matches_operator_chain(sf, threads):
    t = O.q3_generate(sf, threads=2)
    cols, n_semi, n_join = chain(t)
    secs, fp, st = O.bench_q3(t, threads, batch_size=1000)
    assert fp == fingerprint(cols) and fp[0] > 100
    assert st["orders_of_building_customers"] == n_semi and st["joined_rows"] == n_join
    assert st["customer_building"] == int((t["c_mktsegment"] == 1).sum())


def test_q3_generator_matches_pandas_evaluation():
    import pandas as pd
    t = O.q3_generate(0.05)
    assert t["o_orderkey"][:10].tolist() == [1, 2, 3, 4, 5, 6, 7, 8, 33, 34] and t["c_custkey"][0] == 1
    assert t["o_orderdate"].min() >= O.Q3_D0 and t["o_orderdate"].max() <= O.Q3_D1 and set(np.unique(t["c_mktsegment"])) == {0, 1, 2, 3, 4}
    ck = set(t["c_custkey"][t["c_mktsegment"] == 1].tolist())
    od = pd.DataFrame({k: t[k] for k in ("o_orderkey", "o_custkey", "o_orderdate", "o_shippriority")}); od = od[(od.o_orderdate < O.Q3_CUT) & od.o_custkey.isin(ck)]
    ld = pd.DataFrame({k: t[k] for k in ("l_orderkey", "l_extendedprice", "l_discount", "l_shipdate")}); ld = ld[ld.l_shipdate > O.Q3_CUT]
    j = ld.merge(od, left_on="l_orderkey", right_on="o_orderkey")
    j["rev"] = j.l_extendedprice * (100 - j.l_discount)
    g = j.groupby(["l_orderkey", "o_orderdate", "o_shippriority"], as_index=False)["rev"].sum()
    _, fp, _ = O.bench_q3(t, 4)
    assert fp == fingerprint([g.l_orderkey.values, g.o_orderdate.values, g.o_shippriority.values, g.rev.values])


@pytest.mark.parametrize("sf", [0.05, 0.4])
def test_streaming_verifier_agrees_with_the_partitioned_port(sf):
    """the low-memory verifier used for the multi-GPU scale factors (regenerates rows on the fly, shared CAS table) and the
    partitioned port are different algorithms over the same generators: same fingerprint"""
    t = O.q3_generate(sf, threads=2)
    _, fp, st = O.bench_q3(t, 4)
    sfp, joined, qualified = O.q3_stream_fingerprint(sf, threads=3)
    assert sfp == fp and joined == st["joined_rows"] and qualified == st["orders_of_building_customers"]
