"""Run one C2 sparse join + one C3-shaped group-by, device resident (for ncu captures)."""
import sys
sys.path.insert(0, ".")
from datafusion_b200 import capi as D
which = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ctx = D.Context(0)
def col(buf, n): return D.DeviceColumn(ctx, D.INT64, n, buf)
if which in ("all", "join"):
    nb, npr = int(10_000_000 * scale), int(100_000_000 * scale)
    bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, 0, nb); pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb, 0, npr)
    bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, 0, nb); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, 0, npr)
    for it in range(2):
        j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
        j.push_build_device([col(bk, nb), col(bp, nb)]); j.finish_build()
        j.push_probe_device([col(pk, npr), col(pp, npr)]); j.finish_probe()
        print("join rows", j.metric("output_rows"))
        for b in j.drain(host=False): b.release()
        j.close()
if which in ("all", "agg"):
    n, g = int(250_000_000 * scale), 1_000_000
    k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, n); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, n)
    for it in range(2):
        a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
        a.push_device([col(k, n), col(v, n)]); a.finish()
        print("groups", a.metric("num_groups"))
        for b in a.drain(host=False): b.release()
        a.close()
if which in ("all", "filter"):
    n = int(100_000_000 * scale)
    x = ctx.generate_i64(D.GEN_UNIFORM, 1, 0, 1 << 32, 0, n); y = ctx.generate_i64(D.GEN_SPLITMIX, 2, 0, 0, 0, n)
    c = int((1 << 32) * 0.8)   # selectivity 20 % (the reference's default guess, filter.rs:79)
    nodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, c, 0.0), (D.EXPR_BINARY, D.OP_GT, 0, 0, 0, 0.0)]
    for it in range(2):
        f = D.FilterHandle(ctx, [D.INT64, D.INT64], nodes, batch_size=0)
        f.push_device([col(x, n), col(y, n)]); f.finish()
        outs = f.drain(host=False)
        print("filter kept", sum(o.num_rows for o in outs))
        for b in outs: b.release()
        f.close()
ctx.sync()
