"""two C2 joins (100M x 10M, sparse unique keys) with ordered_output = 0 -> the radix-partitioned probe; for ncu captures"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_b200 import capi as D
ctx = D.Context(0)
nb, npr = 10_000_000, 100_000_000
bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, 0, nb); bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, 0, nb)
pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb, 0, npr); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, 0, npr)
col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
for ordered in (False, True, False):
    j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1], ordered_output=ordered)
    j.push_build_device([col(bk, nb), col(bp, nb)]); j.finish_build()
    j.push_probe_device([col(pk, npr), col(pp, npr)]); j.finish_probe()
    print(ordered, j.metric("output_rows"), j.metric("radix_partitioned_probes"))
    for b in j.drain(host=False):
        b.release()
    j.close()
