"""GROUP BY keys beyond the exact 128-bit tag (several columns wider than 128 bits together, a Decimal128 next to other columns): the table
is keyed by a 64-bit hash of the tuple, the claiming thread stores the tuple, and every row is compared against its group's stored tuple
afterwards — GroupValuesColumn's scheme (hash, then vectorized_equal_to, group_values/multi_group_by/mod.rs:628).  Results must equal the
oracle's multi-column group-by."""
import random

import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import batches_to_cols, gpu_group_by

pytestmark = pytest.mark.gpu


def as_dict(cols, nkeys, dec_cols=()):
    """result columns -> {key tuple (None for NULL): tuple of aggregate values}"""
    n = len(cols[0][0])
    out = {}
    for i in range(n):
        key = []
        for c in range(nkeys):
            v, val = cols[c]
            if val is not None and not val[i]:
                key.append(None)
            elif c in dec_cols:
                key.append(D.words_to_decimal(np.asarray(v[i:i + 1]))[0])
            else:
                key.append(np.asarray(v)[i].item())
        vals = []
        for c in range(nkeys, len(cols)):
            v, val = cols[c]
            vals.append(None if (val is not None and not val[i]) else np.asarray(v)[i].item())
        assert tuple(key) not in out, "a group was emitted twice"
        out[tuple(key)] = tuple(vals)
    return out


def oracle_dict(keys, aggs, agg_dtypes):
    ok, res = O.group_by(keys, aggs)
    cols = list(ok)
    for (func, _arg, *_), r, dt in zip(aggs, res, agg_dtypes):
        cols += O.agg_output_columns(func, r, dt, False)
    return as_dict(cols, len(keys))


def wide_inputs(rng, n, n_groups, null_frac):
    universe = [(int(rng.integers(-2**62, 2**62)), int(rng.integers(0, 3)) * 2**40, int(rng.integers(0, 4)), int(rng.integers(-100, 100))) for _ in range(n_groups)]
    pick = rng.integers(0, n_groups, n)
    k = [np.array([universe[p][c] for p in pick], dt) for c, dt in enumerate((np.int64, np.int64, np.int32, np.int64))]
    nul = lambda: None if null_frac == 0 else rng.random(n) >= null_frac
    keys = [(k[0], nul()), (k[1], None), (k[2], nul()), (k[3], nul())]
    v = rng.integers(-10**6, 10**6, n).astype(np.int64)
    vv = rng.random(n) > 0.1
    return keys, (v, vv)


@pytest.mark.parametrize("n_groups,null_frac,batch_rows", [(500, 0.0, None), (40_000, 0.05, 20_000), (3, 0.3, 7_000)])
def test_wide_group_key_vs_oracle(gpu_ctx, n_groups, null_frac, batch_rows):
    """(int64, int64, int32, int64) = 224 bits; several batches, table growth from the default size, NULLs as group values"""
    rng = np.random.default_rng(n_groups)
    keys, val = wide_inputs(rng, 120_000, n_groups, null_frac)
    want = oracle_dict(keys, [(O.A_SUM, val, None), (O.A_COUNT, val, None), (O.A_COUNT_STAR, None, None), (O.A_MIN, val, None)], [np.int64] * 4)
    got = gpu_group_by(gpu_ctx, keys + [val], [0, 1, 2, 3], [(D.AGG_SUM, 4, -1), (D.AGG_COUNT, 4, -1), (D.AGG_COUNT_STAR, -1, -1), (D.AGG_MIN, 4, -1)], batch_rows=batch_rows)
    assert as_dict(got, 4) == want and len(want) >= min(n_groups, 3)


def test_wide_group_key_partial_then_final(gpu_ctx):
    rng = np.random.default_rng(77)
    keys, val = wide_inputs(rng, 90_000, 9_000, 0.04)
    want = oracle_dict(keys, [(O.A_SUM, val, None), (O.A_AVG, val, None)], [np.int64, np.float64])
    cols = keys + [val]
    types = [D.INT64, D.INT64, D.INT32, D.INT64, D.INT64]
    aggs = [(D.AGG_SUM, 4, -1), (D.AGG_AVG, 4, -1)]
    states = []
    for lo, hi in ((0, 50_000), (50_000, 90_000)):
        part = [(c[0][lo:hi], None if c[1] is None else c[1][lo:hi]) for c in cols]
        states.append(gpu_group_by(gpu_ctx, part, [0, 1, 2, 3], aggs, mode=D.AGG_PARTIAL, types=types, batch_rows=17_000))
    st_types = [D.INT64, D.INT64, D.INT32, D.INT64, D.INT64, D.UINT64, D.FLOAT64]
    merged = [(np.concatenate([np.asarray(s[c][0]) for s in states]),
               None if all(s[c][1] is None for s in states) else np.concatenate([np.ones(len(s[c][0]), bool) if s[c][1] is None else s[c][1] for s in states]))
              for c in range(7)]
    got = gpu_group_by(gpu_ctx, merged, [0, 1, 2, 3], aggs, mode=D.AGG_FINAL, types=st_types)
    g, w = as_dict(got, 4), want
    assert g.keys() == w.keys()
    for k in w:
        assert g[k][0] == w[k][0]
        assert (g[k][1] is None and w[k][1] is None) or abs(g[k][1] - w[k][1]) <= 1e-9 * max(1.0, abs(w[k][1]))


def test_decimal_key_next_to_other_group_columns(gpu_ctx):
    """GROUP BY (Decimal128(38,0), int32, float64 with -0.0 / +0.0): 128 + 32 + 64 bits; the oracle groups on (low word, high word, ...)"""
    r = random.Random(5)
    n = 60_000
    pool = [-1, 0, 1, 10**30, -(10**30), (1 << 64), (1 << 64) + 1] + [r.randint(-10**37, 10**37) for _ in range(30)]
    dk = [pool[r.randrange(len(pool))] for _ in range(n)]
    dvalid = np.array([r.random() > 0.04 for _ in range(n)], bool)
    ik = np.array([r.randrange(3) for _ in range(n)], np.int32)
    fk = np.array([[0.0, -0.0, 1.5, float("nan")][r.randrange(4)] for _ in range(n)], np.float64)
    v = np.array([r.randint(-1000, 1000) for _ in range(n)], np.int64)
    words = D.decimal_to_words(dk)
    # oracle: the decimal as two int64 key columns; floats grouped with -0.0 folded into +0.0 and NaN == NaN (primitive.rs:75-98)
    okeys = [(words[:, 0].view(np.int64).copy(), dvalid), (words[:, 1].view(np.int64).copy(), dvalid), (ik, None), (fk, None)]
    ok, res = O.group_by(okeys, [(O.A_SUM, (v, None), None), (O.A_COUNT_STAR, None, None)])
    want = {}
    for i in range(len(ok[0][0])):
        isnull = ok[0][1] is not None and not ok[0][1][i]
        u = ((int(ok[1][0][i]) % 2**64) << 64) | (int(ok[0][0][i]) % 2**64)
        dec = None if isnull else (u - 2**128 if u >= 2**127 else u)
        f = float(ok[3][0][i])
        want[(dec, int(ok[2][0][i]), "nan" if f != f else f + 0.0)] = (int(res[0]["i"][i]), int(res[1]["c"][i]))
    h = D.AggHandle(gpu_ctx, [D.decimal128(38, 0), D.INT32, D.FLOAT64, D.INT64], [0, 1, 2], [(D.AGG_SUM, 3, -1), (D.AGG_COUNT_STAR, -1, -1)], D.AGG_SINGLE)
    for s in range(0, n, 25_000):
        e = min(n, s + 25_000)
        h.push_host([D.HostColumn(words[s:e], dvalid[s:e], D.decimal128(38, 0)), D.HostColumn(ik[s:e]), D.HostColumn(fk[s:e]), D.HostColumn(v[s:e])])
    h.finish()
    got = {}
    for b in h.drain(host=True):
        assert b.column(0).type == D.decimal128(38, 0)
        kd, kdv = b.column_numpy(0)
        ki, kf = b.column_numpy(1)[0], b.column_numpy(2)[0]
        s1, c1 = b.column_numpy(3)[0], b.column_numpy(4)[0]
        decs = D.words_to_decimal(kd)
        for i in range(b.num_rows):
            f = float(kf[i])
            key = (None if (kdv is not None and not kdv[i]) else decs[i], int(ki[i]), "nan" if f != f else f + 0.0)
            assert key not in got
            got[key] = (int(s1[i]), int(c1[i]))
    h.close()
    assert got == want and len(want) > 100
