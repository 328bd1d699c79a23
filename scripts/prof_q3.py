"""one warm-up + one measured run of the device-resident Q3 pipeline (for `ncu --metrics gpu__time_duration.sum` launch lists)"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from datafusion_b200 import capi as D
import q3_device_pipeline as Q
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
ctx = D.Context(0)
t = Q.gen_tables(ctx, sf)
for it in range(2):
    res, st = Q.run_q3(ctx, *t)
    for b in res: b.release()
    ctx.sync()
    print("MARK run", it, st, flush=True)
