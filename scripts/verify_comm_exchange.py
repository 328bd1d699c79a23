"""N processes, one per GPU, NO torch.distributed / NCCL: the exchange control lives in the library (dfgpu_comm / dfgpu_exchange).
python scripts/verify_comm_exchange.py [N]: the parent makes a unique id, spawns N ranks; every rank hash-exchanges its shard of both
join inputs (dfgpu_exchange_run), joins locally, and the union of the rank results must equal the oracle's global join."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, uid, q):
    import numpy as np
    from datafusion_b200 import capi as D
    ctx = D.Context(rank)
    comm = D.Comm(ctx, world, rank, uid)
    assert comm.allgather_i64([rank * 10 + 1, 7]) == [[r * 10 + 1, 7] for r in range(world)]
    nb, npr = 200_003, 2_000_017
    bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, rank * nb, nb); bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, rank * nb, nb)
    pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb * world, rank * npr, npr); pp = ctx.generate_i64(D.GEN_SEQ, 0, rank * 10**10, 0, 0, npr)
    col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
    xb = D.Exchange(comm, [D.INT64, D.INT64], int(nb * 1.5)); xp = D.Exchange(comm, [D.INT64, D.INT64], int(npr * 1.5))
    tot = None
    for rep in range(2):          # twice: the persistent buffers are reused
        eb = xb.run([col(bk, nb), col(bp, nb)], [0]); ep = xp.run([col(pk, npr), col(pp, npr)], [0])
        j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
        j.push_build_device(eb); j.finish_build(); j.push_probe_device(ep); j.finish_probe()
        outs = j.drain(host=True)
        loc = np.stack([np.concatenate([o.column_numpy(c)[0] for o in outs]) for c in range(3)], axis=1) if outs else np.zeros((0, 3), np.int64)
        j.close()
        rows = comm.allgather_i64([len(loc), int(loc.view(np.uint64).sum(dtype=np.uint64) % (2**62)), xb.rows, xp.rows])
        tot = (sum(r[0] for r in rows), sum(r[1] for r in rows) % (2**62), sum(r[2] for r in rows), sum(r[3] for r in rows))
    if rank == 0:
        q.put(tot)
    comm.barrier()
    xb.close(); xp.close(); comm.close(); ctx.close()


def main(world):
    import numpy as np
    from datafusion_b200 import capi as D
    from oracle import oracle as O
    uid = D.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, p.exitcode
    nb, npr = 200_003, 2_000_017
    gbk = O.generate_i64(2, 42, 0, 0, nb * world); gbp = O.generate_i64(2, 7, 0, 0, nb * world)
    gpk = O.generate_i64(4, 42, 43, nb * world, npr * world)
    gpp = np.concatenate([np.arange(npr, dtype=np.int64) + r * 10**10 for r in range(world)])
    exp = O.hash_join([(gbk, None), (gbp, None)], [(gpk, None), (gpp, None)], [0], [0], [0, 0, 1], [0, 1, 1], phj_threshold=0, phj_density=float("inf"))
    em = np.stack([e[0] for e in exp], axis=1)
    want = (len(em), int(em.view(np.uint64).sum(dtype=np.uint64) % (2**62)), nb * world, npr * world)
    assert got == want, (got, want)
    print(f"VERIFY_COMM_EXCHANGE OK N={world} rows={got[0]}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
