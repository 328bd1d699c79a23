"""First end-to-end GPU sanity run (numpy cross-check + rough timings). Not a test, not the bench."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from datafusion_b200 import capi as D

ctx = D.Context(0)
rng = np.random.default_rng(1)

def check_join(nb, npr, dup=1, null_frac=0.0):
    bk = rng.permutation(nb // dup * 3)[: nb // dup].astype(np.int64)
    bk = np.repeat(bk, dup); rng.shuffle(bk)
    bp = rng.integers(0, 1 << 40, nb).astype(np.int64)
    pk = rng.integers(0, nb // dup * 3, npr).astype(np.int64)
    pp = rng.integers(0, 1 << 40, npr).astype(np.int64)
    bvalid = None if null_frac == 0 else rng.random(nb) >= null_frac
    j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1, 1], [0, 1, 0, 1])
    j.push_build_host([D.HostColumn(bk, bvalid), D.HostColumn(bp)])
    j.finish_build()
    j.push_probe_host([D.HostColumn(pk), D.HostColumn(pp)])
    j.finish_probe()
    outs = j.drain(host=True)
    got = [np.concatenate([o.column_numpy(c)[0] for o in outs]) if outs else np.zeros(0, np.int64) for c in range(4)]
    # oracle (numpy): for each probe row in order, matching build rows ascending
    order = np.argsort(bk, kind="stable")
    sk = bk[order]
    ok = np.ones(nb, bool) if bvalid is None else bvalid
    lo = np.searchsorted(sk, pk, "left"); hi = np.searchsorted(sk, pk, "right")
    exp_b, exp_p = [], []
    for i in range(npr):
        rows = order[lo[i]:hi[i]]
        rows = rows[ok[rows]]
        rows = np.sort(rows)
        exp_b.extend(rows.tolist()); exp_p.extend([i] * len(rows))
    exp_b = np.array(exp_b, np.int64); exp_p = np.array(exp_p, np.int64)
    assert len(got[0]) == len(exp_b), (len(got[0]), len(exp_b))
    assert np.array_equal(got[0], bk[exp_b]) and np.array_equal(got[1], bp[exp_b])
    assert np.array_equal(got[2], pk[exp_p]) and np.array_equal(got[3], pp[exp_p])
    print("join ok", nb, npr, dup, null_frac, "rows", len(exp_b), "unique", j.metric("build_unique"), "amap", j.metric("array_map_created_count"))

def check_agg(n, g, null_frac=0.0):
    k = (rng.integers(0, g, n) * 7919 - 5).astype(np.int64)
    v = rng.integers(-2**31, 2**31, n).astype(np.int64)
    valid = None if null_frac == 0 else rng.random(n) >= null_frac
    a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)])
    a.push_host([D.HostColumn(k), D.HostColumn(v, valid)])
    a.finish()
    outs = a.drain(host=True)
    gk = np.concatenate([o.column_numpy(0)[0] for o in outs]); gs = np.concatenate([o.column_numpy(1)[0] for o in outs]); gc = np.concatenate([o.column_numpy(2)[0] for o in outs])
    o = np.argsort(gk); gk, gs, gc = gk[o], gs[o], gc[o]
    uk, inv = np.unique(k, return_inverse=True)
    vv = v if valid is None else np.where(valid, v, 0)
    es = np.zeros(len(uk), np.int64); np.add.at(es, inv, vv)
    ec = np.bincount(inv, weights=None if valid is None else valid.astype(np.int64), minlength=len(uk)).astype(np.int64)
    assert np.array_equal(gk, uk) and np.array_equal(gs, es) and np.array_equal(gc, ec), "agg mismatch"
    print("agg ok", n, g, null_frac, "groups", len(uk), "rehashes", a.metric("rehashes"), "cap", a.metric("table_capacity"))

def check_filter(n, sel):
    x = rng.integers(0, 1 << 32, n).astype(np.int64)
    y = rng.integers(0, 100, n).astype(np.int32)
    c = int(np.quantile(x, 1 - sel))
    nodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, c, 0.0), (D.EXPR_BINARY, D.OP_GT, 0, 0, 0, 0.0)]
    f = D.FilterHandle(ctx, [D.INT64, D.INT32], nodes, batch_size=8192)
    for s in range(0, n, 8192):
        f.push_host([D.HostColumn(x[s:s + 8192]), D.HostColumn(y[s:s + 8192])])
    f.finish()
    outs = f.drain(host=True)
    gx = np.concatenate([o.column_numpy(0)[0] for o in outs]); gy = np.concatenate([o.column_numpy(1)[0] for o in outs])
    m = x > c
    assert np.array_equal(gx, x[m]) and np.array_equal(gy, y[m]), "filter mismatch"
    print("filter ok", n, sel, "kept", int(m.sum()), "batches", len(outs))

check_join(1000, 5000)
check_join(1000, 5000, dup=4)
check_join(50000, 200000, dup=1, null_frac=0.1)
check_join(30000, 100000, dup=3)
check_agg(100000, 1000)
check_agg(2000000, 300000, null_frac=0.05)
check_filter(1 << 20, 0.2)
check_filter(100000, 0.99)

# ---- rough timings, device resident ----
def time_join(nb, npr, kind):
    if kind == "dense":
        bk = ctx.generate_i64(D.GEN_PERM, 42, 0, nb, 0, nb); pk = ctx.generate_i64(D.GEN_UNIFORM, 43, 0, nb, 0, npr)
    else:
        bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, 0, nb); pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb, 0, npr)
    bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, 0, nb); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, 0, npr)
    def col(buf, n): return D.DeviceColumn(ctx, D.INT64, n, buf)
    times = []
    for it in range(4):
        e0, e1, e2 = ctx.event(), ctx.event(), ctx.event()
        j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
        ctx.record(e0)
        j.push_build_device([col(bk, nb), col(bp, nb)]); j.finish_build()
        ctx.record(e1)
        j.push_probe_device([col(pk, npr), col(pp, npr)]); j.finish_probe()
        ctx.record(e2)
        tb, tp = ctx.elapsed_ms(e0, e1), ctx.elapsed_ms(e1, e2)
        rows = j.metric("output_rows")
        times.append((tb, tp))
        for b in j.drain(host=False): b.release()
        j.close()
    print(f"join {kind} {nb}x{npr}: out_rows={rows} build/probe ms: {times}")

def time_agg(n, g):
    k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, n); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, n)
    times = []
    for it in range(4):
        e0, e1 = ctx.event(), ctx.event()
        a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
        ctx.record(e0)
        a.push_device([D.DeviceColumn(ctx, D.INT64, n, k), D.DeviceColumn(ctx, D.INT64, n, v)]); a.finish()
        ctx.record(e1)
        times.append(ctx.elapsed_ms(e0, e1))
        ng = a.metric("num_groups")
        for b in a.drain(host=False): b.release()
        a.close()
    print(f"agg {n} rows {g} groups: ngroups={ng} ms: {times}")

time_join(10_000_000, 100_000_000, "dense")
time_join(10_000_000, 100_000_000, "sparse")
time_agg(250_000_000, 1_000_000)
time_agg(1_000_000_000, 1_000_000)
print("launches", ctx.launches)
