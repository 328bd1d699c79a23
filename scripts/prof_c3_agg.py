"""C3-shaped group-by for ncu: one run per DFGPU_AGG_PAIRED mode given on the command line (python scripts/prof_c3_agg.py ROWS MODE...)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_b200 import capi as D
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
modes = sys.argv[2:] or ["0", "1"]
ctx = D.Context(0)
g = 1_000_000
k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, rows); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, rows)
for m in modes:
    os.environ["DFGPU_AGG_PAIRED"] = m
    a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
    a.push_device([D.DeviceColumn(ctx, D.INT64, rows, k), D.DeviceColumn(ctx, D.INT64, rows, v)]); a.finish()
    res = a.drain(host=False)
    print("mode", m, "groups", sum(b.num_rows for b in res))
    for b in res: b.release()
    a.close()
