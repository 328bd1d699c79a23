"""BASELINE config C5: 8 x B200 radix-partitioned HashJoinExec, 10B x 1B rows — every rank owns 1.25B probe and 125M build rows
{k:int64, payload:int64}; both sides are hash-partitioned on the GPU and exchanged by peer-memory scatter over NVLink (PartitionedHashJoin),
then joined locally.  torchrun --nproc-per-node N scripts/c5_join.py [probe_rows_per_gpu] [build_rows_per_gpu] [steps]
The output is verified at full size: rows == probe rows (100 % hit, unique build keys) and sum(k + 3 pb + 5 pp) mod 2^64 against
independent generator passes on the device (no join involved)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from datafusion_b200 import capi as D, exchange

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
npr = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_250_000_000
nb = int(float(sys.argv[2])) if len(sys.argv) > 2 else 125_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
M64 = (1 << 64) - 1
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pj = exchange.PartitionedHashJoin(local, dist, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1], int(nb * 1.1), int(npr * 1.05), n_chunks=1,
                                  ordered_output=False)
ctx = pj.ctx
torch.cuda.set_stream(pj.js)
NB = nb * world
bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, rank * nb, nb); bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, rank * nb, nb)
pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, NB, rank * npr, npr); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, rank * npr, npr)
col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
build_cols, probe_cols = [col(bk, nb), col(bp, nb)], [col(pk, npr), col(pp, npr)]
# expected fingerprint from independent generator passes: sum k = sum pk, sum pp, sum pb = sum over probe rows of splitmix(7, j_i)
pbx = ctx.generate_i64(D.GEN_SPARSE_OF, 7, 43, NB, rank * npr, npr)
exp_local = (D.column_sum_device(ctx, probe_cols[0]) + 3 * D.column_sum_device(ctx, col(pbx, npr)) + 5 * D.column_sum_device(ctx, probe_cols[1])) & M64
pbx.free()


def allsum(vals):
    t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    return [sum((int(a[i].item()) & M64) for a in allt) & M64 for i in range(len(vals))]


def step(keep):
    return pj.run(build_cols, probe_cols, keep_output=keep)


for _ in range(2):
    step(False)
ctx.sync(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = ctx.event(), ctx.event()
ctx.record(e0)
for _ in range(steps):
    step(False)
ctx.record(e1)
ms = ctx.elapsed_ms(e0, e1) / steps
t = torch.tensor([ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
rows, outs = step(True)
s = [0, 0, 0]
for b in outs:
    for c in range(3):
        s[c] = (s[c] + D.column_sum_device(ctx, b.column(c))) & M64
got = allsum([rows, (s[0] + 3 * s[1] + 5 * s[2]) & M64])
exp = allsum([npr, exp_local])
if rank == 0:
    assert got == exp, f"C5 fingerprint {got} != {exp}"
    total = (nb + npr) * world
    print(json.dumps({"config": f"C5 {world} x B200 partitioned HashJoinExec, {npr * world} x {nb * world} rows int64 ({npr} x {nb} per GPU), sparse unique keys, 100% hit",
                      "n_gpus": world, "ms_per_step": ms, "rows_per_s": total / ms * 1e3, "output_rows": got[0], "fingerprint_verified": True,
                      "exchange_bytes_per_gpu": 16 * (nb + npr) * (world - 1) / world, "nvlink_gbs_per_gpu_lower_bound": 16 * (nb + npr) * (world - 1) / world / ms / 1e6}))
for b in outs:
    b.release()
dist.barrier()
dist.destroy_process_group()
