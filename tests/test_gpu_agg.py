"""Parity of the CUDA AggregateExec path with the reference (golden KATs + oracle).  Group order is
unspecified (the reference's is first-seen per partition): rows are compared sorted, as the
reference's own aggregate fuzzers do (aggregation_fuzzer/mod.rs:59-86).  Integers: bit-exact;
float SUM/AVG: 1e-9 relative (accumulation order differs)."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import assert_cols_equal, col_from_list, gpu_group_by, load_golden

pytestmark = pytest.mark.gpu
MISC = load_golden("misc_kat.json")
F = {O.A_SUM: D.AGG_SUM, O.A_COUNT: D.AGG_COUNT, O.A_MIN: D.AGG_MIN, O.A_MAX: D.AGG_MAX, O.A_AVG: D.AGG_AVG, O.A_COUNT_STAR: D.AGG_COUNT_STAR}


def oracle_table(keys, aggs, arg_dtypes, state=False, **kw):
    ok, res = O.group_by(keys, aggs, **kw)
    cols = list(ok)
    for (func, *_), r, dt in zip(aggs, res, arg_dtypes):
        cols += O.agg_output_columns(func, r, dt, state)
    return cols


def close_cols(got, exp, float_cols):
    g = [c for i, c in enumerate(got) if i not in float_cols]; e = [c for i, c in enumerate(exp) if i not in float_cols]
    order_g = np.lexsort([np.asarray(c[0]).astype(np.float64) for c in g[::-1]]) if g else None
    order_e = np.lexsort([np.asarray(c[0]).astype(np.float64) for c in e[::-1]]) if e else None
    assert_cols_equal(g, e, ordered=False)
    for i in float_cols:
        a, b = np.asarray(got[i][0])[order_g], np.asarray(exp[i][0])[order_e]
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9), f"float column {i}"


def test_gpu_aggregate_some_data(gpu_ctx):
    m = MISC["aggregate_some_data"]
    a = np.concatenate([np.array(b["a"], np.uint32) for b in m["batches"]]); v = np.concatenate([np.array(b["b"], np.float64) for b in m["batches"]])
    part = gpu_group_by(gpu_ctx, [(a, None), (v, None)], [0], [(D.AGG_AVG, 1, -1)], mode=D.AGG_PARTIAL, batch_rows=4)
    o = np.argsort(part[0][0])
    assert part[0][0][o].tolist() == m["partial"]["a"] and part[1][0][o].tolist() == m["partial"]["count"] and part[2][0][o].tolist() == m["partial"]["sum"]
    # Final over two identical partitions of partial state (check_aggregates, aggregates/mod.rs:3660-3700)
    st = [(np.concatenate([c[0], c[0]]), None) for c in part]
    fin = gpu_group_by(gpu_ctx, st, [0], [(D.AGG_AVG, -1, -1)], mode=D.AGG_FINAL, types=[D.UINT32, D.UINT64, D.FLOAT64])
    o = np.argsort(fin[0][0])
    assert fin[0][0][o].tolist() == m["final_avg"]["a"] and fin[1][0][o].tolist() == m["final_avg"]["avg"]


def test_gpu_sum_count_null_state_and_wrapping(gpu_ctx):
    m = MISC["sum_null_state"]
    g = np.array(m["g"] + m["second_batch"]["g"], np.int64)
    v = col_from_list(m["v"] + m["second_batch"]["v"], np.int64)
    got = gpu_group_by(gpu_ctx, [(g, None), v], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], batch_rows=len(m["g"]))
    o = np.argsort(got[0][0])
    exp = m["expected"]
    assert got[0][0][o].tolist() == exp["g"]
    sums = [None if (got[1][1] is not None and not got[1][1][i]) else int(got[1][0][i]) for i in o]
    assert sums == exp["sum"] and got[2][0][o].tolist() == exp["count"] and got[2][1] is None


@pytest.mark.parametrize("n,groups,null_frac", [(50_000, 300, 0.0), (200_000, 50_000, 0.07), (1_000_000, 400_000, 0.0)])
def test_gpu_vs_oracle_single_key(gpu_ctx, n, groups, null_frac):
    rng = np.random.default_rng(n + groups)
    k = (rng.integers(0, groups, n) * 104729 - 17).astype(np.int64)
    kv = None if null_frac == 0 else rng.random(n) >= null_frac / 2       # NULL group keys form one group (primitive.rs:144-148)
    v = rng.integers(-2**62, 2**62, n).astype(np.int64)                    # sums wrap (sum.rs:316)
    vv = None if null_frac == 0 else rng.random(n) >= null_frac
    f = rng.standard_normal(n)
    filt = (rng.random(n) > 0.3, None if null_frac == 0 else rng.random(n) > 0.1)
    cols = [(k, kv), (v, vv), (f, vv), filt]
    aggs = [(O.A_SUM, (v, vv), None), (O.A_COUNT, (v, vv), None), (O.A_MIN, (v, vv), None), (O.A_MAX, (v, vv), None), (O.A_AVG, (f, vv), None),
            (O.A_SUM, (v, vv), filt), (O.A_COUNT_STAR, None, None)]
    exp = oracle_table([(k, kv)], aggs, [np.int64, np.int64, np.int64, np.int64, np.float64, np.int64, np.int64])
    gaggs = [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1), (D.AGG_MIN, 1, -1), (D.AGG_MAX, 1, -1), (D.AGG_AVG, 2, -1), (D.AGG_SUM, 1, 3), (D.AGG_COUNT_STAR, -1, -1)]
    for batch_rows, device in ((8192, False), (None, True)):
        got = gpu_group_by(gpu_ctx, cols, [0], gaggs, batch_rows=batch_rows, device=device)
        close_cols(got, exp, float_cols=[5])


def test_gpu_multi_column_keys_and_128bit(gpu_ctx):
    # TPC-H Q3 group key shape: (l_orderkey int64, o_orderdate date32, o_shippriority int32) = 128 bits
    rng = np.random.default_rng(21)
    n = 300_000
    k1 = rng.integers(-2**62, 2**62, 20000).astype(np.int64)[rng.integers(0, 20000, n)]
    k2 = rng.integers(8000, 8010, n).astype(np.int32); k3 = rng.integers(0, 2, n).astype(np.int32)
    v = rng.integers(0, 10**9, n).astype(np.int64)
    exp = oracle_table([(k1, None), (k2, None), (k3, None)], [(O.A_SUM, (v, None), None), (O.A_COUNT, (v, None), None)], [np.int64, np.int64])
    got, h = gpu_group_by(gpu_ctx, [(k1, None), (k2, None), (k3, None), (v, None)], [0, 1, 2], [(D.AGG_SUM, 3, -1), (D.AGG_COUNT, 3, -1)],
                          types=[D.INT64, D.DATE32, D.INT32, D.INT64], return_handle=True)
    assert h.metric("key_words") == 2
    h.close()
    assert_cols_equal(got, exp, ordered=False)
    # nullable two-column key (null flags live in the key): (NULL, 1) and (NULL, 2) are different groups
    a = col_from_list([1, None, None, 1, None, 2], np.int32); b = col_from_list([1, 1, 2, 1, 1, None], np.int32)
    vv = (np.arange(6, dtype=np.int64), None)
    exp = oracle_table([a, b], [(O.A_SUM, vv, None)], [np.int64])
    got = gpu_group_by(gpu_ctx, [a, b, vv], [0, 1], [(D.AGG_SUM, 2, -1)])
    assert_cols_equal(got, exp, ordered=False)


def test_gpu_partial_then_final_equals_single(gpu_ctx):
    # Partial -> (exchange) -> Final merge must equal Single (aggregates/mod.rs:28-48)
    rng = np.random.default_rng(8)
    n = 400_000
    k = rng.integers(0, 30_000, n).astype(np.int64); v = rng.integers(-10**12, 10**12, n).astype(np.int64); vv = rng.random(n) > 0.05
    cols = [(k, None), (v, vv)]
    aggs = [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1), (D.AGG_MIN, 1, -1), (D.AGG_MAX, 1, -1)]
    single = gpu_group_by(gpu_ctx, cols, [0], aggs)
    parts = [gpu_group_by(gpu_ctx, [(k[s:e], None), (v[s:e], vv[s:e])], [0], aggs, mode=D.AGG_PARTIAL) for s, e in ((0, 150_000), (150_000, 400_000))]
    merged_in = []
    for c in range(5):
        vals = np.concatenate([p[c][0] for p in parts])
        valid = np.concatenate([np.ones(len(p[c][0]), bool) if p[c][1] is None else p[c][1] for p in parts])
        merged_in.append((vals, None if valid.all() else valid))
    final = gpu_group_by(gpu_ctx, merged_in, [0], [(f, -1, -1) for f, _, _ in aggs], mode=D.AGG_FINAL)
    assert_cols_equal(final, single, ordered=False)


def test_gpu_table_growth_and_float_keys(gpu_ctx):
    rng = np.random.default_rng(4)
    n = 1_500_000
    k = rng.permutation(n).astype(np.int64)            # every row its own group: forces overflow replay + rehash from 64K slots
    v = np.ones(n, np.int64)
    got, h = gpu_group_by(gpu_ctx, [(k, None), (v, None)], [0], [(D.AGG_SUM, 1, -1)], return_handle=True)
    assert h.metric("num_groups") == n and h.metric("rehashes") >= 1
    h.close()
    assert np.array_equal(np.sort(got[0][0]), np.arange(n)) and (got[1][0] == 1).all()
    fk = np.array([0.0, -0.0, 1.5, np.nan, np.nan, -0.0], np.float64)   # -0.0 folds into +0.0; NaN groups by bits (primitive.rs:75-98)
    exp = oracle_table([(fk, None)], [(O.A_COUNT_STAR, None, None)], [np.int64])
    got = gpu_group_by(gpu_ctx, [(fk, None)], [0], [(D.AGG_COUNT_STAR, -1, -1)])
    assert_cols_equal(got, exp, ordered=False)


def test_gpu_large_groupby_properties(gpu_ctx):
    """BASELINE config C3 scaled (100M rows here, 1M groups; bench.py/scripts run 1B): checksum of the
    group table against the oracle's Partial->Final multi-threaded run over the same generators."""
    ctx = gpu_ctx
    n, g = 100_000_000, 1_000_000
    k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, n); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, n)
    a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
    a.push_device([D.DeviceColumn(ctx, D.INT64, n, k), D.DeviceColumn(ctx, D.INT64, n, v)]); a.finish()
    outs = a.drain(host=True)
    gk = np.concatenate([o.column_numpy(0)[0] for o in outs]); gs = np.concatenate([o.column_numpy(1)[0] for o in outs]); gc = np.concatenate([o.column_numpy(2)[0] for o in outs])
    assert len(gk) == g and int(gc.sum()) == n and len(np.unique(gk)) == g
    hk = O.generate_i64(1, 5, 0, g, n, 8); hv = O.generate_i64(1, 6, -2**31, 2**32, n, 8)
    secs, groups, chk = O.bench_groupby(hk, hv, threads=8)
    mine = int(gk.view(np.uint64).sum(dtype=np.uint64) * np.uint64(3) + gs.view(np.uint64).sum(dtype=np.uint64) * np.uint64(5) + gc.view(np.uint64).sum(dtype=np.uint64) * np.uint64(7))
    assert groups == g and mine % 2**64 == chk
    a.close()


def test_gpu_single_aggregate_planning_kat(gpu_ctx):
    m = MISC["single_aggregate_planning"]
    got = gpu_group_by(gpu_ctx, [(np.array(m["a_u32"], np.uint32), None), (np.array(m["b_f64"], np.float64), None)], [0], [(D.AGG_SUM, 1, -1)], batch_rows=2)
    o = np.argsort(got[0][0])
    assert got[0][0][o].tolist() == m["expected"]["a"] and got[1][0][o].tolist() == m["expected"]["sum"]


@pytest.mark.parametrize("name", ["skip_aggregation_after_first_batch", "skip_aggregation_after_threshold"])
def test_gpu_final_accepts_the_references_skip_aggregation_states(gpu_ctx, name):
    """aggregates/mod.rs:5431-5603: the reference's Partial may pass rows through unaggregated; GPU Final over exactly those
    states == GPU Single over the raw rows == the reference's final counts."""
    m = MISC[name]
    key = np.concatenate([np.array(b["key"], np.int32) for b in m["batches"]]); val = np.concatenate([np.array(b["val"], np.int32) for b in m["batches"]])
    single = gpu_group_by(gpu_ctx, [(key, None), (val, None)], [0], [(D.AGG_COUNT, 1, -1)], batch_rows=3)
    rp = m["reference_partial"]
    final = gpu_group_by(gpu_ctx, [(np.array(rp["key"], np.int32), None), (np.array(rp["count"], np.int64), None)], [0], [(D.AGG_COUNT, -1, -1)], mode=D.AGG_FINAL)
    for got in (single, final):
        o = np.argsort(got[0][0])
        assert got[0][0][o].tolist() == m["final"]["key"] and got[1][0][o].tolist() == m["final"]["count"]
    # and our own Partial states (always aggregated) merge to the same result
    part = gpu_group_by(gpu_ctx, [(key, None), (val, None)], [0], [(D.AGG_COUNT, 1, -1)], mode=D.AGG_PARTIAL, batch_rows=3)
    fin2 = gpu_group_by(gpu_ctx, part, [0], [(D.AGG_COUNT, -1, -1)], mode=D.AGG_FINAL)
    o = np.argsort(fin2[0][0])
    assert fin2[0][0][o].tolist() == m["final"]["key"] and fin2[1][0][o].tolist() == m["final"]["count"]


def test_gpu_vectorized_group_values_intern_kat(gpu_ctx):
    """multi_group_by/mod.rs:1986-1997, 2260-2540 (VectorizedTestDataSet): three nullable key columns with every NULL / repeated /
    already-in-map combination over three batches -> the reference's 17 groups (string columns dictionary-coded to int16)"""
    from test_oracle_golden import group_rows, vectorized_group_values_case
    cols, exp, sizes = vectorized_group_values_case()
    cols = [cols[0], (cols[1][0].astype(np.int16), cols[1][1]), (cols[2][0].astype(np.int16), cols[2][1])]
    ones = (np.ones(len(cols[0][0]), np.int64), None)
    for batch_rows in (14, 5, None):
        got = gpu_group_by(gpu_ctx, cols + [ones], [0, 1, 2], [(D.AGG_COUNT, 3, -1)], batch_rows=batch_rows)
        assert group_rows(got[:3]) == exp
        assert int(got[3][0].sum()) == sum(sizes)


def test_gpu_check_grouping_sets_kat(gpu_ctx):
    """aggregates/mod.rs:3428-3590 (check_grouping_sets): GROUPING SETS ((a), (b), (a,b)) are expressed through the ABI as one pass per
    set with the masked-out group columns pushed as all-NULL columns plus the UInt8 __grouping_id key — Single, and Partial -> Final."""
    from test_oracle_golden import grouping_rows, grouping_sets_case
    keys, ones, exp = grouping_sets_case()
    cols = keys + [ones]
    got = gpu_group_by(gpu_ctx, cols, [0, 1, 2], [(D.AGG_COUNT, 3, -1)], batch_rows=8)      # one push per (set, input batch) of the reference
    assert grouping_rows(got[:3], got[3][0]) == exp
    part = gpu_group_by(gpu_ctx, cols, [0, 1, 2], [(D.AGG_COUNT, 3, -1)], mode=D.AGG_PARTIAL, batch_rows=8)
    fin = gpu_group_by(gpu_ctx, part, [0, 1, 2], [(D.AGG_COUNT, -1, -1)], mode=D.AGG_FINAL)
    assert grouping_rows(fin[:3], fin[3][0]) == exp


@pytest.mark.parametrize("n,batch_rows", [(0, None), (1, None), (300_000, 70_001)])
def test_gpu_aggregate_without_group_by(gpu_ctx, n, batch_rows):
    """AggregateStream (no GROUP BY, aggregate_stream.rs): one output row, also for empty input; Single == Final(Partial) for every function"""
    rng = np.random.default_rng(5 + n)
    v = rng.integers(-10**12, 10**12, n).astype(np.int64); vv = rng.random(n) > 0.1
    f = rng.normal(size=n); u = rng.integers(0, 2**40, n).astype(np.uint64)
    flt = rng.random(n) > 0.5
    cols = [(v, vv if n else None), (f, None), (u, None), (flt, None)]
    aggs = [(D.AGG_SUM, 0, -1), (D.AGG_COUNT, 0, -1), (D.AGG_MIN, 0, 3), (D.AGG_MAX, 2, -1), (D.AGG_AVG, 1, -1), (D.AGG_COUNT_STAR, -1, -1), (D.AGG_SUM, 2, 3)]
    oaggs = [(O.A_SUM, cols[0], None), (O.A_COUNT, cols[0], None), (O.A_MIN, cols[0], cols[3]), (O.A_MAX, cols[2], None), (O.A_AVG, cols[1], None),
             (O.A_COUNT_STAR, None, None, n), (O.A_SUM, cols[2], cols[3])]
    exp = O.scalar_aggregate(oaggs)
    got = gpu_group_by(gpu_ctx, cols, [], aggs, batch_rows=batch_rows, device=True)
    assert len(got) == len(exp) and all(len(c[0]) == 1 for c in got)
    for i, (g, e) in enumerate(zip(got, exp)):
        gvalid = g[1] is None or bool(g[1][0]); evalid = e[1] is None or bool(e[1][0])
        assert gvalid == evalid, f"aggregate {i} validity"
        if evalid:
            if np.asarray(e[0]).dtype.kind == "f":
                assert np.isclose(float(g[0][0]), float(e[0][0]), rtol=1e-9, atol=1e-9), f"aggregate {i}"
            else:
                assert int(g[0][0]) == int(e[0][0]), f"aggregate {i}"
    # Partial -> Final gives the same row
    part = gpu_group_by(gpu_ctx, cols, [], aggs, mode=D.AGG_PARTIAL, batch_rows=batch_rows, device=True)
    pexp = O.scalar_aggregate(oaggs, state=True)
    assert len(part) == len(pexp) == 8
    types = [D.INT64, D.INT64, D.INT64, D.UINT64, D.UINT64, D.FLOAT64, D.INT64, D.UINT64]
    st = [(np.concatenate([c[0], c[0][:0]]), c[1]) for c in part]
    fin = gpu_group_by(gpu_ctx, st, [], [(a[0], -1, -1) for a in aggs], mode=D.AGG_FINAL, types=types)
    for i, (g, e) in enumerate(zip(fin, exp)):
        gvalid = g[1] is None or bool(g[1][0]); evalid = e[1] is None or bool(e[1][0])
        assert gvalid == evalid
        if evalid:
            assert np.isclose(float(g[0][0]), float(e[0][0]), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name,threshold,n_agg", [("skip_aggregation_after_first_batch", 2, 3), ("skip_aggregation_after_threshold", 5, 4)])
def test_gpu_skip_partial_reproduces_the_references_partial_output(gpu_ctx, name, threshold, n_agg):
    """aggregates/mod.rs:5431-5603 with the reference's settings (probe_rows_threshold 2 / 5, ratio 0.1): the GPU Partial emits the
    aggregated groups, then every later batch row by row (convert_to_state) in input order — the reference's snapshot."""
    m = MISC[name]
    key = np.concatenate([np.array(b["key"], np.int32) for b in m["batches"]]); val = np.concatenate([np.array(b["val"], np.int32) for b in m["batches"]])
    got, h = gpu_group_by(gpu_ctx, [(key, None), (val, None)], [0], [(D.AGG_COUNT, 1, -1)], mode=D.AGG_PARTIAL, batch_rows=3, skip_partial=(threshold, 0.1), return_handle=True)
    rp = m["reference_partial"]
    assert h.metric("skipped_aggregation_rows") == len(rp["key"]) - n_agg
    h.close()
    assert sorted(zip(got[0][0][:n_agg].tolist(), got[1][0][:n_agg].tolist())) == sorted(zip(rp["key"][:n_agg], rp["count"][:n_agg]))
    assert got[0][0][n_agg:].tolist() == rp["key"][n_agg:] and got[1][0][n_agg:].tolist() == rp["count"][n_agg:]
    # the probe is off with threshold 0, and the default thresholds (100 000 rows, 0.8) never trigger here
    for sp in ((0, 0.1), None):
        got = gpu_group_by(gpu_ctx, [(key, None), (val, None)], [0], [(D.AGG_COUNT, 1, -1)], mode=D.AGG_PARTIAL, batch_rows=3, skip_partial=sp)
        assert sorted(zip(got[0][0].tolist(), got[1][0].tolist())) == sorted(zip(m["final"]["key"], m["final"]["count"]))


def test_gpu_skip_partial_high_cardinality_final_equals_single(gpu_ctx):
    """default thresholds: 300 000 rows with ~all-distinct keys -> after the first 120 000-row batch groups / rows > 0.8 and the rest
    passes through as state rows (SUM / COUNT / MIN / AVG / COUNT(*), NULL values, a FILTER, NULL keys); Final over that == Single"""
    rng = np.random.default_rng(101)
    n = 300_000
    k = rng.permutation(10 * n)[:n].astype(np.int64); kv = rng.random(n) > 0.01
    v = rng.integers(-10**9, 10**9, n).astype(np.int64); vv = rng.random(n) > 0.1
    f = rng.normal(size=n); flt = rng.random(n) > 0.3
    cols = [(k, kv), (v, vv), (f, None), (flt, None)]
    aggs = [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, 3), (D.AGG_MIN, 1, -1), (D.AGG_AVG, 2, -1), (D.AGG_COUNT_STAR, -1, -1), (D.AGG_MAX, 2, 3)]
    part, h = gpu_group_by(gpu_ctx, cols, [0], aggs, mode=D.AGG_PARTIAL, batch_rows=120_000, device=True, return_handle=True)
    assert h.metric("skipped_aggregation_rows") == n - 120_000 and len(part[0][0]) > 0.95 * n
    h.close()
    types = [D.INT64, D.INT64, D.INT64, D.INT64, D.UINT64, D.FLOAT64, D.INT64, D.FLOAT64]
    fin = gpu_group_by(gpu_ctx, part, [0], [(a[0], -1, -1) for a in aggs], mode=D.AGG_FINAL, types=types)
    single = gpu_group_by(gpu_ctx, cols, [0], aggs, batch_rows=120_000)
    close_cols(fin, single, float_cols={4, 6})
    oaggs = [(O.A_SUM, cols[1], None), (O.A_COUNT, cols[1], cols[3]), (O.A_MIN, cols[1], None), (O.A_AVG, cols[2], None), (O.A_COUNT_STAR, None, None), (O.A_MAX, cols[2], cols[3])]
    exp = oracle_table([cols[0]], oaggs, [np.int64, np.int64, np.int64, np.float64, np.int64, np.float64])
    close_cols(single, exp, float_cols={4, 6})
