"""one fused Q3 run (for ncu): python scripts/prof_q3_fused.py SF"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from datafusion_b200 import capi as D
import q3_device_pipeline as Q
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = D.Context(0)
cu, orr, li = Q.gen_tables(ctx, sf)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    res, st = Q.run_q3_fused(ctx, cu, orr, li)
    for b in res: b.release()
print(st)
