import sys
sys.path.insert(0, ".")
from datafusion_b200 import capi as D
ctx = D.Context(0)
n, g = 1_000_000_000, 1_000_000
k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, n); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, n)
ctx.set_kernel_timing(True)
times = []
for it in range(4):
    e0, e1 = ctx.event(), ctx.event()
    a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
    ctx.record(e0)
    a.push_device([D.DeviceColumn(ctx, D.INT64, n, k), D.DeviceColumn(ctx, D.INT64, n, v)]); a.finish()
    ctx.record(e1)
    times.append(round(ctx.elapsed_ms(e0, e1), 3))
    for b in a.drain(host=False): b.release()
    a.close()
print("agg 1B rows 1M groups ms:", times, "kernel:", ctx.kernel_time("agg_update"))
