"""per-phase CUDA-event timing of one C2 join step (diagnostics)"""
import sys
sys.path.insert(0, ".")
from datafusion_b200 import capi as D
ctx = D.Context(0)
nb, npr = 10_000_000, 100_000_000
kind = sys.argv[1] if len(sys.argv) > 1 else "sparse"
if kind == "dense":
    bk = ctx.generate_i64(D.GEN_PERM, 42, 0, nb, 0, nb); pk = ctx.generate_i64(D.GEN_UNIFORM, 43, 0, nb, 0, npr)
else:
    bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, 0, nb); pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb, 0, npr)
bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, 0, nb); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, 0, npr)
col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
ctx.set_kernel_timing(True)
for it in range(5):
    ev = [ctx.event() for _ in range(6)]
    ctx.record(ev[0])
    j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
    j.push_build_device([col(bk, nb), col(bp, nb)]); ctx.record(ev[1])
    j.finish_build(); ctx.record(ev[2])
    j.push_probe_device([col(pk, npr), col(pp, npr)]); ctx.record(ev[3])
    j.finish_probe(); outs = j.drain(host=False); ctx.record(ev[4])
    for b in outs: b.release()
    j.close(); ctx.record(ev[5])
    t = [ctx.elapsed_ms(ev[i], ev[i + 1]) for i in range(5)]
    print(kind, "push_build %.3f finish_build %.3f push_probe %.3f drain %.3f release %.3f | total %.3f" % (*t, ctx.elapsed_ms(ev[0], ev[5])), "amap", j.metric("array_map_created_count") if False else "")
print("probe kernel", ctx.kernel_time("join_probe"), "build kernel", ctx.kernel_time("join_build"))
