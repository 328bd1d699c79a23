"""torchrun --nproc-per-node N scripts/verify_q3_multi_gpu.py [SF per rank]: the partitioned multi-GPU Q3 plan (scripts/q3_multi_gpu.py)
must give — summed over ranks — the result fingerprint of an independent CPU evaluation of the same SF(sf x N) database
(oracle_q3_stream_fingerprint) and, at N == 1 world sizes, of the single-GPU fused pipelines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import q3_multi_gpu as QM
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
r = QM.PartitionedQ3(local, dist, sf)
for it in range(2):                      # twice: the persistent filter / exchange buffers are reused correctly
    st = r.step()
    fp = r.all_reduce_fingerprint(r.fingerprint())
    if rank == 0:
        from oracle import oracle as O
        efp, ejoined, eorders = O.q3_stream_fingerprint(sf * world, seed=1, threads=min(8, os.cpu_count() or 1))
        assert fp[:5] == efp and fp[5] == ejoined and fp[6] == eorders, f"step {it}: {fp} != {efp + [ejoined, eorders]}"
        print(f"step {it}: N={world} SF{sf:g}/rank fingerprint ok {fp[:5]} joined {fp[5]} stages(rank0) {st}")
dist.barrier()
if rank == 0:
    print("VERIFY_Q3_MULTI_GPU OK")
dist.destroy_process_group()
