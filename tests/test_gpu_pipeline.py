"""Parity of the fused pipeline operator (dfgpu_lookup / dfgpu_pipeline: FilterExec -> HashJoinExec probe side(s) -> sink in one
kernel) with the UNFUSED operator chain evaluated by the oracle (filter_batch -> hash_join -> eval_expr -> group_by, each pinned
by the reference's own vectors in tests/test_oracle_golden.py).  Integer results are bit-exact; group order is unspecified
(compared sorted, as the reference's aggregate fuzzers do); the output sink preserves input order like RightSemi / Inner with a
unique build side do (hash_join/exec.rs:634-640)."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import assert_cols_equal, batches_to_cols, host_cols, split_points
from test_gpu_filter import B, C, L, to_nodes

pytestmark = pytest.mark.gpu
F = {D.AGG_SUM: O.A_SUM, D.AGG_COUNT: O.A_COUNT, D.AGG_MIN: O.A_MIN, D.AGG_MAX: O.A_MAX, D.AGG_AVG: O.A_AVG, D.AGG_COUNT_STAR: O.A_COUNT_STAR}


def push_all(p, cols, types, batch_rows, device, ctx, keep):
    n = len(cols[0][0])
    for s, e in split_points(n, batch_rows):
        hc = host_cols(cols, s, e, types)
        if device:
            dc = [D.DeviceColumn.from_host(ctx, h) for h in hc]
            keep.append(dc)
            p.push_device(dc)
        else:
            p.push_host(hc)


def build_lookup(ctx, cols, types, key_col, payload_cols, pred=None, stages=(), batch_rows=None, device=False, **lk):
    look = D.Lookup(ctx, types[key_col], [types[c] for c in payload_cols] if not stages else lk.pop("payload_types"), **lk)
    p = D.Pipeline(ctx, types, to_nodes(pred, True) if pred is not None else None, stages)
    p.sink_build(look, key_col, payload_cols)
    keep = []
    push_all(p, cols, types, batch_rows, device, ctx, keep)
    p.finish()
    rows = p.metric("sink_rows")
    p.close()
    return look, rows


def oracle_filter(cols, pred):
    if pred is None:
        return list(cols)
    return O.filter_batch(cols, O.eval_expr(cols, to_nodes(pred, False)))


@pytest.mark.parametrize("mode", ["bitmap", "hash", "hash_filter"])
@pytest.mark.parametrize("kind", [D.STAGE_SEMI, D.STAGE_ANTI])
def test_pipeline_semi_anti_output(gpu_ctx, mode, kind):
    rng = np.random.default_rng(3)
    nb, npr = 20_000, 150_000
    bk = rng.integers(-500, 60_000, nb).astype(np.int64)            # duplicates on the build side are fine for a key set
    bseg = rng.integers(0, 5, nb).astype(np.int32)
    pk = rng.integers(-2_000, 80_000, npr).astype(np.int64); pkv = rng.random(npr) > 0.03
    pd = rng.integers(0, 1000, npr).astype(np.int32); pv = rng.integers(-10**12, 10**12, npr).astype(np.int64)
    build, probe = [(bk, None), (bseg, None)], [(pk, pkv), (pd, None), (pv, None)]
    bpred, ppred = B(D.OP_EQ, C(1), L(1, np.int32)), B(D.OP_AND, B(D.OP_LT, C(1), L(700, np.int32)), B(D.OP_GTEQ, C(2), L(-5 * 10**11, np.int64)))
    kr = D.column_minmax_device(gpu_ctx, D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(bk))) if mode == "bitmap" else None
    if kr is not None:
        assert kr == (int(bk.min()), int(bk.max()), nb)
    look, rows = build_lookup(gpu_ctx, build, [D.INT64, D.INT32], 0, [], pred=bpred, key_range=None if kr is None else kr[:2],
                              membership_filter=1 if mode == "hash_filter" else 0, batch_rows=7000)
    assert look.metric("mode") == (1 if mode == "bitmap" else 0)
    assert rows == int((bseg == 1).sum())
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT32, D.INT64], to_nodes(ppred, True), [(kind, 0, look)])
    p.sink_output([2, 1], batch_size=8192)      # the nullable key column itself stays out of the fused output (non-null columns only)
    keep = []
    push_all(p, probe, None, 40_000, True, gpu_ctx, keep)
    p.finish()
    outs = p.drain(host=False)
    assert all(o.num_rows == 8192 for o in outs[:-1])
    got = batches_to_cols(outs, 2)
    fb, fp = oracle_filter(build, bpred), oracle_filter(probe, ppred)
    # RightSemi keeps probe rows with a partner, RightAnti those without — NULL probe keys never match, so Anti emits them (utils.rs:1461-1476)
    exp = O.hash_join(fb, fp, [0], [0], [1, 1], [2, 1], join_type=O.J_RIGHT_SEMI if kind == D.STAGE_SEMI else O.J_RIGHT_ANTI)
    assert 0 < len(exp[0][0]) < len(fp[0][0])
    assert_cols_equal(got, exp, ordered=True, what=f"semi/anti {mode}")
    p.close(); look.close()


def test_pipeline_inner_output_with_payload_fields(gpu_ctx):
    rng = np.random.default_rng(5)
    nb, npr = 30_000, 200_000
    bk = (rng.permutation(200_000)[:nb].astype(np.int64) - 1000) * 3          # unique, sparse, some negative
    bd = rng.integers(8000, 10000, nb).astype(np.int32); bp = rng.integers(0, 3, nb).astype(np.int16)
    pk = ((rng.integers(0, 200_000, npr) - 1000) * 3).astype(np.int64)
    px = rng.integers(0, 100, npr).astype(np.int32)
    build, probe = [(bk, None), (bd, None), (bp, None)], [(px, None), (pk, None)]
    look, rows = build_lookup(gpu_ctx, build, [D.INT64, D.DATE32, D.INT16], 0, [1, 2], device=True, batch_rows=11_000)
    assert rows == nb and look.metric("rows") == nb and look.metric("rehashes") >= 1   # no expected_rows: the table grew by rehash
    p = D.Pipeline(gpu_ctx, [D.INT32, D.INT64], None, [(D.STAGE_INNER, 1, look)])
    p.sink_output([1, 0, 2, 3])                                               # probe key, probe x, build date, build prio
    keep = []
    push_all(p, probe, [D.INT32, D.INT64], 64_000, False, gpu_ctx, keep)
    p.finish()
    got = batches_to_cols(p.drain(host=True), 4)
    exp = O.hash_join(build, probe, [0], [1], [1, 1, 0, 0], [1, 0, 1, 2])
    assert_cols_equal(got, exp, ordered=True, what="inner output")
    p.close(); look.close()


def q3_like_tables(rng, nc, no, nl, nulls):
    c_key = np.arange(1, nc + 1, dtype=np.int64); c_seg = rng.integers(0, 5, nc).astype(np.int64)
    o_key = (np.arange(no, dtype=np.int64) // 8) * 32 + (np.arange(no, dtype=np.int64) % 8) + 1
    o_cust = rng.integers(1, nc * 2 // 3 + 2, no).astype(np.int64)
    o_date = rng.integers(8000, 10500, no).astype(np.int32); o_prio = rng.integers(0, 2, no).astype(np.int32)
    l_key = o_key[rng.integers(0, no, nl)]
    l_price = rng.integers(90_000, 10_500_000, nl).astype(np.int64); l_disc = rng.integers(0, 11, nl).astype(np.int64)
    l_ship = rng.integers(8000, 10600, nl).astype(np.int32)
    lk_valid = (rng.random(nl) > 0.02) if nulls else None
    lp_valid = (rng.random(nl) > 0.05) if nulls else None
    oc_valid = (rng.random(no) > 0.02) if nulls else None
    return ([(c_key, None), (c_seg, None)],
            [(o_key, None), (o_cust, oc_valid), (o_date, None), (o_prio, None)],
            [(l_key, lk_valid), (l_price, lp_valid), (l_disc, None), (l_ship, None)])


@pytest.mark.parametrize("nulls,device,batch_rows", [(False, True, None), (False, False, 50_000), (True, True, 33_333)])
def test_pipeline_q3_shape_matches_unfused_oracle_chain(gpu_ctx, nulls, device, batch_rows):
    """the reference's Q3 physical plan (tpch/plans/q3.slt.part:60-76) as three fused pipelines vs the oracle's operator chain"""
    rng = np.random.default_rng(11)
    cust, orders, line = q3_like_tables(rng, 3000, 30_000, 120_000, nulls)
    CUT = 9200
    ct, ot, lt = [D.INT64, D.INT64], [D.INT64, D.INT64, D.DATE32, D.INT32], [D.INT64, D.INT64, D.INT64, D.DATE32]
    cpred, opred, lpred = B(D.OP_EQ, C(1), L(1, np.int64)), B(D.OP_LT, C(2), L(CUT, np.int32)), B(D.OP_GT, C(3), L(CUT, np.int32))
    rev = B(D.OP_MULTIPLY, C(1), B(D.OP_MINUS, L(100, np.int64), C(2)))
    l1, _ = build_lookup(gpu_ctx, cust, ct, 0, [], pred=cpred, key_range=(1, 3000), device=device)
    l2, n2 = build_lookup(gpu_ctx, orders, ot, 0, [2, 3], pred=opred, stages=[(D.STAGE_SEMI, 1, l1)], payload_types=[D.DATE32, D.INT32],
                          n_acc_words=5, membership_filter=1, device=device, batch_rows=batch_rows)
    p = D.Pipeline(gpu_ctx, lt, to_nodes(lpred, True), [(D.STAGE_INNER, 0, l2)])
    aggs = [(D.AGG_SUM, to_nodes(rev, True)), (D.AGG_COUNT_STAR, None), (D.AGG_COUNT, to_nodes(C(1), True)), (D.AGG_MAX, to_nodes(C(2), True))]
    p.sink_aggregate([0, 4, 5], aggs, D.AGG_SINGLE)
    keep = []
    push_all(p, line, lt, batch_rows, device, gpu_ctx, keep)
    p.finish()
    got = batches_to_cols(p.drain(host=not device), 7)
    # ---- the unfused chain on the oracle ----
    fc, fo, fl = oracle_filter(cust, cpred), oracle_filter(orders, opred), oracle_filter(line, lpred)
    so = O.hash_join(fc, fo, [0], [1], [1, 1, 1], [0, 2, 3], join_type=O.J_RIGHT_SEMI)
    assert len(so[0][0]) == n2
    j = O.hash_join(so, fl, [0], [0], [1, 0, 0, 1, 1], [0, 1, 2, 1, 2])
    arg = O.eval_expr([j[0], j[3], j[4]], to_nodes(rev, False))
    keys, res = O.group_by([j[0], j[1], j[2]], [(O.A_SUM, arg, None), (O.A_COUNT_STAR, None, None), (O.A_COUNT, j[3], None), (O.A_MAX, j[4], None)])
    exp = list(keys) + O.agg_output_columns(O.A_SUM, res[0], np.int64, False) + O.agg_output_columns(O.A_COUNT_STAR, res[1], np.int64, False) + \
        O.agg_output_columns(O.A_COUNT, res[2], np.int64, False) + O.agg_output_columns(O.A_MAX, res[3], np.int64, False)
    assert p.metric("num_groups") == len(keys[0][0]) > 1000
    assert_cols_equal(got, exp, ordered=False, what="q3 shape")
    p.close(); l2.close(); l1.close()


def test_pipeline_interpreter_predicate_avg_min_partial_states(gpu_ctx):
    rng = np.random.default_rng(17)
    nb, npr = 5000, 80_000
    bk = rng.permutation(20_000)[:nb].astype(np.int32); bt = rng.integers(0, 7, nb).astype(np.int8)
    pk = rng.integers(0, 20_000, npr).astype(np.int32)
    a = rng.integers(-1000, 1000, npr).astype(np.int64); b = rng.integers(-1000, 1000, npr).astype(np.int64)
    f = rng.normal(size=npr); fv = rng.random(npr) > 0.1
    look, _ = build_lookup(gpu_ctx, [(bk, None), (bt, None)], [D.INT32, D.INT8], 0, [1], n_acc_words=8, expected_rows=nb)
    assert look.metric("rehashes") == 0
    pred = B(D.OP_OR, B(D.OP_GT, B(D.OP_PLUS, C(1), C(2)), L(100, np.int64)), B(D.OP_LT, C(1), L(-900, np.int64)))
    p = D.Pipeline(gpu_ctx, [D.INT32, D.INT64, D.INT64, D.FLOAT64], to_nodes(pred, True), [(D.STAGE_INNER, 0, look)])
    aggs = [(D.AGG_AVG, to_nodes(C(3), True)), (D.AGG_MIN, to_nodes(C(2), True)), (D.AGG_SUM, to_nodes(C(3), True))]
    p.sink_aggregate([4, 0], aggs, D.AGG_PARTIAL)
    probe = [(pk, None), (a, None), (b, None), (f, fv)]
    push_all(p, probe, None, 30_000, True, gpu_ctx, [])
    p.finish()
    got = batches_to_cols(p.drain(host=False), 6)
    fp = oracle_filter(probe, pred)
    j = O.hash_join([(bk, None), (bt, None)], fp, [0], [0], [0, 1, 1, 1], [1, 0, 2, 3])
    keys, res = O.group_by([j[0], j[1]], [(O.A_AVG, j[3], None), (O.A_MIN, j[2], None), (O.A_SUM, j[3], None)])
    exp = list(keys) + O.agg_output_columns(O.A_AVG, res[0], np.float64, True) + O.agg_output_columns(O.A_MIN, res[1], np.int64, False) + \
        O.agg_output_columns(O.A_SUM, res[2], np.float64, False)
    # integer / key / count columns bit-exact; float sums within 1e-9 relative (accumulation order differs)
    og, oe = np.lexsort((got[1][0], got[0][0])), np.lexsort((exp[1][0], exp[0][0]))
    for c in (0, 1, 2, 4):
        assert np.array_equal(np.asarray(got[c][0])[og], np.asarray(exp[c][0])[oe]), f"column {c}"
    for c in (3, 5):
        gv, ev = np.asarray(got[c][0])[og], np.asarray(exp[c][0])[oe]
        gval = np.ones(len(gv), bool) if got[c][1] is None else got[c][1][og]
        eval_ = np.ones(len(ev), bool) if exp[c][1] is None else exp[c][1][oe]
        if c == 3:   # AVG partial sum state: the oracle reports 0.0 for groups without a value; NULL-ness follows the count column
            assert np.array_equal(gval, np.asarray(got[2][0])[og] > 0)
        else:
            assert np.array_equal(gval, eval_)
        assert np.allclose(np.where(gval, gv, 0.0), np.where(gval, ev, 0.0), rtol=1e-9, atol=1e-9), f"float column {c}"
    p.close(); look.close()


def test_pipeline_rejects_what_it_cannot_fuse(gpu_ctx):
    k = np.array([1, 2, 2, 3], np.int64); v = np.array([5, 6, 7, 8], np.int32)
    look = D.Lookup(gpu_ctx, D.INT64, [D.INT32])
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT32])
    p.sink_build(look, 0, [1])
    with pytest.raises(D.DfgpuError) as ei:
        p.push_host([D.HostColumn(k), D.HostColumn(v)])
    assert ei.value.code == -3 and "duplicate" in str(ei.value)
    p.close(); look.close()
    look = D.Lookup(gpu_ctx, D.INT64, [D.INT32], n_acc_words=2)
    b = D.Pipeline(gpu_ctx, [D.INT64, D.INT32]); b.sink_build(look, 0, [1]); b.push_host([D.HostColumn(k[[0, 1, 3]]), D.HostColumn(v[[0, 1, 3]])]); b.finish(); b.close()
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT64], None, [(D.STAGE_INNER, 0, look)])
    with pytest.raises(D.DfgpuError) as ei:     # group key not determined by the join key
        p.sink_aggregate([1], [(D.AGG_COUNT_STAR, None)])
    assert ei.value.code == -3
    with pytest.raises(D.DfgpuError) as ei:     # more accumulators than the lookup reserves
        p.sink_aggregate([0], [(D.AGG_SUM, to_nodes(C(1), True)), (D.AGG_SUM, to_nodes(C(1), True))])
    assert ei.value.code == -3
    p.sink_aggregate([0, 2], [(D.AGG_SUM, to_nodes(C(1), True))])
    with pytest.raises(D.DfgpuError) as ei:     # nullable aggregate input without a spare non-null counter word
        p.push_host([D.HostColumn(np.array([1, 3], np.int64)), D.HostColumn(np.array([4, 4], np.int64), np.array([True, False]))])
    assert ei.value.code == -3
    p.close(); look.close()
    look = D.Lookup(gpu_ctx, D.INT64, [], key_range=(0, 10))
    b = D.Pipeline(gpu_ctx, [D.INT64]); b.sink_build(look, 0, [])
    with pytest.raises(D.DfgpuError) as ei:     # a key outside the promised range
        b.push_host([D.HostColumn(np.array([3, 11], np.int64))])
    assert ei.value.code == -1
    b.close(); look.close()


def test_pipeline_divide_by_zero_in_aggregate_argument_is_an_error(gpu_ctx):
    look = D.Lookup(gpu_ctx, D.INT64, [D.INT32], n_acc_words=2)
    b = D.Pipeline(gpu_ctx, [D.INT64, D.INT32]); b.sink_build(look, 0, [1])
    b.push_host([D.HostColumn(np.array([1, 2], np.int64)), D.HostColumn(np.array([0, 1], np.int32))]); b.finish(); b.close()
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT64], None, [(D.STAGE_INNER, 0, look)])
    p.sink_aggregate([0], [(D.AGG_SUM, to_nodes(B(D.OP_DIVIDE, L(10, np.int64), C(1)), True))])
    with pytest.raises(D.DfgpuError) as ei:
        p.push_host([D.HostColumn(np.array([1, 2, 9], np.int64)), D.HostColumn(np.array([5, 0, 0], np.int64))])
    assert ei.value.code == -4      # row 2 joins and divides by zero; row 3 never reaches the expression (no partner)
    p.close(); look.close()


def test_pipeline_unordered_output_and_maybe_stage(gpu_ctx):
    """the exchange-feeding shape of the multi-GPU plan: scan -> predicate -> MAYBE(membership filter of a downstream join) -> unordered
    output.  The filter has no false negatives (every true partner survives) and few false positives; the exact join downstream gives the
    unfused result."""
    rng = np.random.default_rng(23)
    nb, npr = 40_000, 300_000
    bk = (rng.permutation(400_000)[:nb].astype(np.int64)) * 7 + 3
    pk = (rng.integers(0, 400_000, npr).astype(np.int64)) * 7 + 3
    pv = rng.integers(0, 10**9, npr).astype(np.int64); pd = rng.integers(0, 100, npr).astype(np.int32)
    filt = D.Lookup(gpu_ctx, D.INT64, [], expected_rows=nb, filter_only=True)
    b = D.Pipeline(gpu_ctx, [D.INT64]); b.sink_build(filt, 0, []); b.push_host([D.HostColumn(bk)]); b.finish(); b.close()
    assert filt.filter_buffer()[1] == max(1024, nb // 4) * 8
    pred = B(D.OP_LT, C(2), L(60, np.int32))
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT64, D.INT32], to_nodes(pred, True), [(D.STAGE_MAYBE, 0, filt)])
    p.sink_output([0, 1], ordered=False)
    push_all(p, [(pk, None), (pv, None), (pd, None)], None, 100_000, True, gpu_ctx, [])
    p.finish()
    got = batches_to_cols(p.drain(host=False), 2)
    p.close()
    keep = pd < 60
    true_partner = keep & np.isin(pk, bk)
    gset = set(zip(got[0][0].tolist(), got[1][0].tolist()))
    assert set(zip(pk[true_partner].tolist(), pv[true_partner].tolist())) <= gset          # no false negatives
    assert len(got[0][0]) <= true_partner.sum() + 0.05 * keep.sum()                        # few false positives (16 bits per key)
    assert gset <= set(zip(pk[keep].tolist(), pv[keep].tolist())) and len(gset) == len(got[0][0])
    # downstream exact join over the survivors == the unfused join over everything
    look, _ = build_lookup(gpu_ctx, [(bk, None)], [D.INT64], 0, [], n_acc_words=2, membership_filter=0)
    q = D.Pipeline(gpu_ctx, [D.INT64, D.INT64], None, [(D.STAGE_INNER, 0, look)])
    q.sink_aggregate([0], [(D.AGG_SUM, to_nodes(C(1), True))])
    q.push_host([D.HostColumn(got[0][0]), D.HostColumn(got[1][0])]); q.finish()
    res = batches_to_cols(q.drain(host=True), 2)
    keys, r = O.group_by([(pk[true_partner], None)], [(O.A_SUM, (pv[true_partner], None), None)])
    assert_cols_equal(res, [keys[0]] + O.agg_output_columns(O.A_SUM, r[0], np.int64, False), ordered=False, what="maybe + exact join")
    q.close(); look.close()
    # clear() empties the filter: nothing may pass any more
    filt.clear()
    p = D.Pipeline(gpu_ctx, [D.INT64, D.INT64, D.INT32], None, [(D.STAGE_MAYBE, 0, filt)])
    p.sink_output([0], ordered=False)
    p.push_host([D.HostColumn(pk), D.HostColumn(pv), D.HostColumn(pd)]); p.finish()
    assert p.drain(host=True) == [] and p.metric("sink_rows") == 0
    p.close(); filt.close()


def test_pipeline_unordered_output_matches_ordered(gpu_ctx):
    rng = np.random.default_rng(29)
    n = 500_000
    k = rng.integers(0, 5000, n).astype(np.int64); v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    look = D.Lookup(gpu_ctx, D.INT64, [], key_range=(0, 4999))
    b = D.Pipeline(gpu_ctx, [D.INT64]); b.sink_build(look, 0, []); b.push_host([D.HostColumn(np.arange(0, 5000, 3, dtype=np.int64))]); b.finish(); b.close()
    outs = []
    for ordered in (True, False):
        p = D.Pipeline(gpu_ctx, [D.INT64, D.INT64], to_nodes(B(D.OP_GT, C(1), L(0, np.int64)), True), [(D.STAGE_SEMI, 0, look)])
        p.sink_output([1, 0], ordered=ordered)
        p.push_host([D.HostColumn(k), D.HostColumn(v)]); p.finish()
        outs.append(batches_to_cols(p.drain(host=True), 2))
        p.close()
    m = (v > 0) & (k % 3 == 0)
    assert_cols_equal(outs[0], [(v[m], None), (k[m], None)], ordered=True, what="ordered output")
    assert_cols_equal(outs[1], [(v[m], None), (k[m], None)], ordered=False, what="unordered output")
    look.close()
