"""One process, many answers (GPU minutes are scarce): event-timed A/B of the opt-in kernel variants at the bench sizes.
  * Q3 SF100 fused plan with pipe_kernel's instantiations (DFGPU_PIPE_VAR bits: 1 / 2 prefetches, 8 lane-paired REDs, 32 256-bit column
    loads; read per launch) against the default instantiation (43) and the round-start kernel (0)
  * C3 group-by (1B rows -> 1M groups, SUM + COUNT) with the paired-accumulator kernel's modes (DFGPU_AGG_PAIRED: 4 default, 3 / 4 with the
    256-bit bucket load, 0 = one RED per aggregate and row), read when the handle is created
Every variant's result fingerprint must equal the default's.  Prints one JSON object; `winner` = the fastest variant if it beats the
default by >= 2 %, else the default.  (The two passes of round 2 — profiles/r2c_ab_variants_pass{1,2}.json — are tabulated in profiles/README.md.)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from datafusion_b200 import capi as D
import q3_device_pipeline as Q

Q3_DEFAULT, C3_DEFAULT = 43, 4   # kPipeVarDefault (pipeline.cu), kAggPairedDefault (aggregate.cu)
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000_000
ctx = D.Context(0)
out = {"sf": sf, "agg_rows": rows}


def q3(var, steps=6):
    os.environ["DFGPU_PIPE_VAR"] = str(var)
    for _ in range(2):
        res, st = Q.run_q3_fused(ctx, cu, orr, li)
        fp = Q.result_fingerprint(ctx, res)
        for b in res: b.release()
    ctx.set_kernel_timing(True); ctx.kernel_time_reset()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(steps):
        res, st = Q.run_q3_fused(ctx, cu, orr, li)
        for b in res: b.release()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / steps
    kt = ctx.kernel_time("pipe:lineitem"); ko = ctx.kernel_time("pipe:orders")
    ctx.set_kernel_timing(False)
    return {"var": var, "step_ms": round(ms, 3), "lineitem_kernel_ms": round(kt[0] / max(kt[1], 1), 3), "orders_kernel_ms": round(ko[0] / max(ko[1], 1), 3), "fingerprint": fp}


cu, orr, li = Q.gen_tables(ctx, sf)
runs = [q3(v) for v in (Q3_DEFAULT, 11, 9, 3, 0, Q3_DEFAULT)]
del cu, orr, li
ctx.trim_device_cache()
base = min(r["step_ms"] for r in runs if r["var"] == Q3_DEFAULT)
assert all(r["fingerprint"] == runs[0]["fingerprint"] for r in runs), "a prefetch variant changed the result"
best = min(runs, key=lambda r: r["step_ms"])
out["q3"] = {"runs": runs, "baseline_step_ms": base, "winner": best["var"] if best["step_ms"] < 0.98 * base else Q3_DEFAULT}
print(json.dumps(out), file=sys.stderr, flush=True)   # partial result, in case the second half does not get to run
os.environ["DFGPU_PIPE_VAR"] = str(Q3_DEFAULT)

g = 1_000_000
k = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, g, 0, rows); v = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, rows)


def c3(paired, r4=0, iters=3):
    os.environ["DFGPU_AGG_PAIRED"] = str(paired); os.environ["DFGPU_AGG_R4"] = str(r4)
    times, fp = [], None
    ctx.set_kernel_timing(True); ctx.kernel_time_reset()
    for it in range(iters + 1):
        e0, e1 = ctx.event(), ctx.event()
        a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=g)
        ctx.record(e0)
        a.push_device([D.DeviceColumn(ctx, D.INT64, rows, k), D.DeviceColumn(ctx, D.INT64, rows, v)]); a.finish()
        ctx.record(e1)
        if it: times.append(round(ctx.elapsed_ms(e0, e1), 3))
        res = a.drain(host=False)
        if fp is None:
            ng = sum(b.num_rows for b in res)
            sums = [0, 0, 0]
            for b in res:
                for c in range(3):
                    sums[c] = (sums[c] + D.column_sum_device(ctx, b.column(c))) % (1 << 64)
            fp = [ng] + sums
        for b in res: b.release()
        a.close()
    kt = ctx.kernel_time("agg_update")
    ctx.set_kernel_timing(False)
    return {"paired": paired, "r4": r4, "step_ms": times, "kernel_ms": round(kt[0] / max(kt[1], 1), 3), "fingerprint": fp}


aruns = [c3(C3_DEFAULT), c3(3), c3(1), c3(0), c3(C3_DEFAULT)]
abase = min(min(r["step_ms"]) for r in aruns if r["paired"] == C3_DEFAULT and r["r4"] == 0)
assert all(r["fingerprint"] == aruns[0]["fingerprint"] for r in aruns), "a group-by variant changed the result"
abest = min(aruns, key=lambda r: min(r["step_ms"]))
out["c3"] = {"runs": aruns, "baseline_step_ms": abase,
             "winner_paired": abest["paired"] if min(abest["step_ms"]) < 0.98 * abase else C3_DEFAULT, "winner_r4": 0}
print(json.dumps(out))
