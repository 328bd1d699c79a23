"""Dictionary-coded string keys: per-batch Arrow dictionaries are unified on the host (dfgpu_dictionary), the codes of
each batch are rewritten on the device, and the ordinary integer-key operators run on the codes.  Oracle: pyarrow on the strings."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import batches_to_cols

pytestmark = pytest.mark.gpu
WORDS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD", "", "büilding", "a" * 300]


def string_batches(rng, n_batches, rows, null_frac=0.05, words=WORDS):
    """pyarrow DictionaryArray batches whose dictionaries differ in content and order"""
    import pyarrow as pa
    out = []
    for _ in range(n_batches):
        k = int(rng.integers(1, len(words) + 1))
        local = [words[i] for i in rng.permutation(len(words))[:k]]
        if rng.random() < 0.5:
            local.append(None)                       # a NULL dictionary value: rows pointing at it are NULL
        idx = rng.integers(0, len(local), rows).astype(np.int32)
        mask = rng.random(rows) < null_frac
        arr = pa.DictionaryArray.from_arrays(pa.array(idx, mask=mask), pa.array(local, pa.string()))
        out.append(arr)
    return out


def unified_codes(ctx, dic, arr):
    """one DictionaryArray -> (INT32 device batch of unified codes, python list of the strings)"""
    import pyarrow as pa
    values = arr.dictionary
    bufs = values.buffers()
    offsets = np.frombuffer(bufs[1], np.int32, len(values) + 1, values.offset * 4)
    data = np.frombuffer(bufs[2], np.uint8) if bufs[2] is not None else np.zeros(0, np.uint8)
    valid = None if values.null_count == 0 else np.array(values.is_valid())
    remap = dic.unify(offsets, data, valid)
    idx = arr.indices
    codes = D.HostColumn(np.asarray(idx.fill_null(0)), None if idx.null_count == 0 else np.array(idx.is_valid()))
    return dic.remap(codes, remap, on_host=True), arr.to_pylist()


def test_dictionary_unify_and_lookup(gpu_ctx):
    dic = D.Dictionary(gpu_ctx)
    off = np.array([0, 1, 3, 3], np.int32)
    r1 = dic.unify(off, np.frombuffer(b"abc", np.uint8), None)
    assert r1.tolist() == [0, 1, 2] and dic.size() == 3
    r2 = dic.unify(np.array([0, 2, 2, 3], np.int32), np.frombuffer(b"bca", np.uint8), np.array([True, True, False]))
    assert r2.tolist() == [1, 2, -1] and dic.size() == 3            # "bc" and "" are known; the NULL value maps to -1
    assert dic.code(b"a") == 0 and dic.code(b"") == 2 and dic.code(b"zzz") == -1
    assert [dic.value(i) for i in range(3)] == [b"a", b"bc", b""]
    with pytest.raises(D.DfgpuError, match="outside"):
        dic.remap(D.HostColumn(np.array([0, 7], np.int32)), r1, on_host=True)
    dic.close()


def test_group_by_string_keys_across_batches(gpu_ctx):
    """GROUP BY a dictionary-coded string column (GroupValuesBytes in the reference, group_values/mod.rs:139-217) == pyarrow on the strings"""
    import pyarrow as pa
    rng = np.random.default_rng(5)
    dic = D.Dictionary(gpu_ctx)
    agg = D.AggHandle(gpu_ctx, [D.INT32, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT_STAR, -1, -1)], D.AGG_SINGLE)
    strings, vals, keep = [], [], []
    for arr in string_batches(rng, 6, 20_000):
        b, s = unified_codes(gpu_ctx, dic, arr)
        v = rng.integers(-1000, 1000, len(s)).astype(np.int64)
        vc = D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(v))
        keep.append((b, vc))
        agg.push_device([b.column(0), vc])
        strings += s; vals += v.tolist()
    agg.finish()
    got = batches_to_cols(agg.drain(host=True), 3)
    ref = pa.table({"k": pa.array(strings, pa.string()), "v": vals}).group_by("k").aggregate([("v", "sum"), ([], "count_all")])
    want = {k: (s, c) for k, s, c in zip(ref["k"].to_pylist(), ref["v_sum"].to_pylist(), ref["count_all"].to_pylist())}
    out = {}
    kv, kval = got[0]
    for i in range(len(kv)):
        key = None if (kval is not None and not kval[i]) else dic.value(int(kv[i])).decode()
        out[key] = (int(got[1][0][i]), int(got[2][0][i]))
    assert out == want and len(want) >= 8
    agg.close(); dic.close()


def test_join_on_string_keys_with_separate_dictionaries(gpu_ctx):
    """both sides of a join coded against ONE dictionary: Inner join on the codes == pyarrow join on the strings"""
    import pyarrow as pa
    rng = np.random.default_rng(6)
    dic = D.Dictionary(gpu_ctx)
    build_arr = string_batches(rng, 1, 400, null_frac=0.1)[0]
    probe_arrs = string_batches(rng, 3, 5000, null_frac=0.1)
    j = D.HashJoinHandle(gpu_ctx, [D.INT32, D.INT64], [D.INT32, D.INT64], [0], [0], [0, 1], [1, 1])
    bb, bs = unified_codes(gpu_ctx, dic, build_arr)
    bid = np.arange(len(bs), dtype=np.int64)
    bidc = D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(bid))
    j.push_build_device([bb.column(0), bidc]); j.finish_build()
    ps, pid, keep = [], [], []
    base = 0
    for arr in probe_arrs:
        pb, s = unified_codes(gpu_ctx, dic, arr)
        ids = np.arange(base, base + len(s), dtype=np.int64); base += len(s)
        idc = D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(ids))
        keep.append((pb, idc))
        j.push_probe_device([pb.column(0), idc])
        ps += s; pid += ids.tolist()
    j.finish_probe()
    got = batches_to_cols(j.drain(host=True), 2)
    pairs = sorted(zip(got[0][0].tolist(), got[1][0].tolist()))
    ref = pa.table({"k": pa.array(bs, pa.string()), "b": bid}).join(pa.table({"k": pa.array(ps, pa.string()), "p": pid}), keys="k", join_type="inner")
    want = sorted(zip(ref["b"].to_pylist(), ref["p"].to_pylist()))
    assert pairs == want and len(want) > 1000
    j.close(); dic.close()


def test_filter_on_string_literal(gpu_ctx):
    """c_mktsegment = 'BUILDING' (q3.slt.part:69) as `codes = code('BUILDING')`; a literal that was never seen matches nothing"""
    rng = np.random.default_rng(7)
    dic = D.Dictionary(gpu_ctx)
    arr = string_batches(rng, 1, 30_000)[0]
    b, s = unified_codes(gpu_ctx, dic, arr)
    ids = np.arange(len(s), dtype=np.int64)
    for word in ("BUILDING", "never seen"):
        code = dic.code(word.encode())
        nodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT32, 0, code, 0.0), (D.EXPR_BINARY, D.OP_EQ, 0, 0, 0, 0.0)]
        f = D.FilterHandle(gpu_ctx, [D.INT32, D.INT64], nodes, [1], 8192, -1)
        f.push_device([b.column(0), D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(ids))])
        f.finish()
        got = batches_to_cols(f.drain(host=True), 1)[0][0].tolist()
        assert got == [i for i, x in enumerate(s) if x == word]
        f.close()
    dic.close()
