import sys
sys.path.insert(0, ".")
from datafusion_b200 import capi as D
ctx = D.Context(0)
n = 100_000_000
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
k = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, 0, n); v = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, 0, n)
cols = [D.DeviceColumn(ctx, D.INT64, n, k), D.DeviceColumn(ctx, D.INT64, n, v)]
ctx.set_kernel_timing(True)
for it in range(4):
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    b, offs = D.hash_partition_device(ctx, cols, [0], P)
    ctx.record(e1)
    print("partition", P, "parts: %.3f ms" % ctx.elapsed_ms(e0, e1))
    b.release()
print(ctx.kernel_time("partition"))
