"""torchrun script (N >= 2 GPUs): PartitionedAggregate (Partial -> exchange of states -> FinalPartitioned) must equal the
oracle's single-pass group-by over the concatenated input, and BroadcastHashJoin (CollectLeft analogue) must produce the
same multiset of rows as the oracle's global join."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, torch.distributed as dist
from datafusion_b200 import capi as D, exchange
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
ctx = D.Context(local, ts.cuda_stream)
ok = True


def gather_rows(cols):
    """(values, valid) columns of this rank -> rank 0 gets the concatenation over ranks"""
    objs = [None] * world
    dist.all_gather_object(objs, cols)
    out = []
    for c in range(len(cols)):
        v = np.concatenate([o[c][0] for o in objs])
        val = np.concatenate([np.ones(len(o[c][0]), bool) if o[c][1] is None else o[c][1] for o in objs])
        out.append((v, None if val.all() else val))
    return out


def batch_cols(batches, ncols):
    out = []
    for c in range(ncols):
        vals, valids = [], []
        for b in batches:
            col = b.column(c)
            v = ctx.to_host(col.values, b.num_rows * D.WIDTH[col.type]).view(D.NP_OF_TYPE[col.type]).copy()
            vals.append(v)
            if col.validity:
                bits = np.unpackbits(ctx.to_host(col.validity, (b.num_rows + 7) // 8), bitorder="little")[:b.num_rows].astype(bool)
            else:
                bits = np.ones(b.num_rows, bool)
            valids.append(bits)
        v = np.concatenate(vals) if vals else np.zeros(0, np.int64)
        val = np.concatenate(valids) if valids else np.zeros(0, bool)
        out.append((v, None if val.all() else val))
    return out


from harness import assert_cols_equal
from oracle import oracle as O

# ---- group-by: (a) C3 shape, no NULLs; (b) NULL values + MIN/MAX/AVG, rank-dependent nullability; (c) one rank without rows
cases = []
rng = np.random.default_rng(1000 + rank)
n = 400_000 + 1000 * rank
cases.append(("sum_count", rng.integers(0, 50_000, n).astype(np.int64), (rng.integers(-2**31, 2**31, n).astype(np.int64), None),
              [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)]))
v = rng.integers(-1000, 1000, n).astype(np.int64)
vv = rng.random(n) > (0.5 if rank == 0 else 0.0)        # only rank 0 has NULLs: its partial state is nullable, the others' is not
cases.append(("nulls_minmaxavg", rng.integers(0, 3000, n).astype(np.int64), (v, None if vv.all() else vv),
              [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1), (D.AGG_MIN, 1, -1), (D.AGG_MAX, 1, -1), (D.AGG_AVG, 1, -1)]))
m = 0 if rank == world - 1 else 50_000
cases.append(("empty_rank", rng.integers(0, 100, m).astype(np.int64), (rng.integers(0, 9, m).astype(np.int64), None), [(D.AGG_SUM, 1, -1), (D.AGG_COUNT_STAR, -1, -1)]))
for name, k, (v, vv), aggs in cases:
    kc = D.DeviceColumn.from_host(ctx, D.HostColumn(k)); vc = D.DeviceColumn.from_host(ctx, D.HostColumn(v, vv))
    pa = exchange.PartitionedAggregate(ctx, dist, [D.INT64, D.INT64], [0], aggs, capacity_hint=0)
    outs = pa.run([kc, vc])
    ncols = 1 + len(aggs)
    mine = batch_cols(outs, ncols)
    got = gather_rows(mine)
    allk = gather_rows([(k, None), (v, vv)])
    if rank == 0:
        gk, gv = allk
        OF = {D.AGG_SUM: O.A_SUM, D.AGG_COUNT: O.A_COUNT, D.AGG_MIN: O.A_MIN, D.AGG_MAX: O.A_MAX, D.AGG_AVG: O.A_AVG, D.AGG_COUNT_STAR: O.A_COUNT_STAR}
        oaggs = [(OF[f], (gv if a >= 0 else None), None) for f, a, _ in aggs]
        okeys, ores = O.group_by([gk], oaggs)
        exp = list(okeys)
        for (f, *_), r in zip(oaggs, ores):
            exp += O.agg_output_columns(f, r, np.int64, False)
        try:
            fl = [i for i, c in enumerate(exp) if np.asarray(c[0]).dtype.kind == "f"]
            keep = [i for i in range(len(exp)) if i not in fl]
            assert_cols_equal([got[i] for i in keep], [exp[i] for i in keep], ordered=False, what=name)
            og = np.argsort(got[0][0], kind="stable"); oe = np.argsort(exp[0][0], kind="stable")
            for i in fl:   # AVG over int64: sum in f64; accumulation order differs across partial/final -> 1e-9 relative
                assert np.allclose(np.asarray(got[i][0])[og], np.asarray(exp[i][0])[oe], rtol=1e-9, atol=1e-9), f"float column {i}"
            print(f"group-by {name}: {len(got[0][0])} groups == oracle", flush=True)
        except AssertionError as e:
            ok = False
            print(f"group-by {name}: MISMATCH {str(e)[:400]}", flush=True)
    for b in outs: b.release()

# ---- broadcast (CollectLeft) join: small build side replicated, probe side stays put
nb, npr = 5_000 + rank, 300_000
bk = (np.arange(nb, dtype=np.int64) * world + rank) * 3; bp = rng.integers(0, 1 << 40, nb).astype(np.int64)
pk = rng.integers(0, (5_000 + world) * world, npr).astype(np.int64) * 3; pp = np.arange(npr, dtype=np.int64) + rank * 10**9
bj = exchange.BroadcastHashJoin(ctx, dist, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1, 1], [0, 1, 0, 1])
dc = lambda a: D.DeviceColumn.from_host(ctx, D.HostColumn(a))
rows, outs = bj.run([dc(bk), dc(bp)], [dc(pk), dc(pp)])
got = gather_rows(batch_cols(outs, 4))
allb = gather_rows([(bk, None), (bp, None)]); allp = gather_rows([(pk, None), (pp, None)])
if rank == 0:
    exp = O.hash_join(allb, allp, [0], [0], [0, 0, 1, 1], [0, 1, 0, 1])
    try:
        assert_cols_equal(got, exp, ordered=False, what="broadcast join")
        print(f"broadcast join: {len(got[0][0])} rows == oracle", flush=True)
    except AssertionError as e:
        ok = False
        print(f"broadcast join: MISMATCH {str(e)[:400]}", flush=True)
if rank == 0:
    print("multi_gpu_ops_ok=%s" % ok, flush=True)
dist.barrier(); dist.destroy_process_group()
