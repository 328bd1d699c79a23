"""per-phase timing of the exchange + local join on N GPUs (torchrun)"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from datafusion_b200 import capi as D, exchange
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
ctx = D.Context(local, ts.cuda_stream)
nb, npr = 10_000_000, 100_000_000
bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, rank * nb, nb); bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, rank * nb, nb)
pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb * world, rank * npr, npr); pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, rank * npr, npr)
col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
pc = [col(pk, npr), col(pp, npr)]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    e0 = ev()
    batch, offs = D.hash_partition_device(ctx, pc, [0], world)
    e1 = ev()
    send_counts = [offs[p + 1] - offs[p] for p in range(world)]
    recv_counts = exchange.exchange_counts(dist, send_counts, torch.device("cuda", local))
    e2 = ev()
    views = []
    for i in range(batch.num_columns):
        c = batch.column(i)
        views.append(torch.as_tensor(exchange._CudaView(c.values, c.length, "<i8", batch), device=torch.device("cuda", local)))
    e3 = ev()
    recv = exchange.all_to_all_columns(dist, views, send_counts, recv_counts)
    e4 = ev()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"it{it}: partition {e0.elapsed_time(e1):.3f} counts {e1.elapsed_time(e2):.3f} wrap {e2.elapsed_time(e3):.3f} all_to_all(2 cols) {e3.elapsed_time(e4):.3f} ms; sent {sum(send_counts) - send_counts[rank]} rows off-GPU")
dist.destroy_process_group()
