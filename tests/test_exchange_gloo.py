"""world_size-2 CPU (gloo) test of the N>1 host logic: per-destination counts exchange + all-to-all-v per
column + local join must reproduce the global join (the exchange of RepartitionExec, repartition/mod.rs:1097-1145)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _partition_ids(keys: np.ndarray, n_parts: int) -> np.ndarray:
    from oracle import oracle as O  # noqa: F401  (hash choice is irrelevant to the result; use a simple mixer)
    k = keys.astype(np.uint64)
    h = (k ^ (k >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    h ^= h >> np.uint64(27)
    return (h % np.uint64(n_parts)).astype(np.int64)


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datafusion_b200 import exchange
    from oracle import oracle as O
    rng = np.random.default_rng(100 + rank)
    nb, npr = 4000, 30000
    # rank-local shards of both tables (global keys overlap across ranks)
    bk = (np.arange(nb, dtype=np.int64) * world + rank) * 7 - 3
    bp = rng.integers(0, 1 << 40, nb).astype(np.int64)
    pk = rng.integers(0, nb * world, npr).astype(np.int64) * 7 - 3
    pp = np.arange(npr, dtype=np.int64) + rank * 10**9

    def exchange_cpu(cols, key):
        pid = _partition_ids(cols[key], world)
        order = np.argsort(pid, kind="stable")             # stable: rows keep input order inside each partition
        send_counts = np.bincount(pid, minlength=world).tolist()
        recv_counts = exchange.exchange_counts(dist, send_counts, torch.device("cpu"))
        sc, rc, so, ro = exchange.plan_all_to_all(send_counts, recv_counts)
        assert so[-1] == len(cols[0]) and ro[-1] == sum(recv_counts)
        sent = [torch.from_numpy(np.ascontiguousarray(c[order])) for c in cols]
        got = exchange.all_to_all_columns(dist, sent, sc, rc)
        return [g.numpy() for g in got], rc

    (rbk, rbp), _ = exchange_cpu([bk, bp], 0)
    (rpk, rpp), rc = exchange_cpu([pk, pp], 0)
    assert (_partition_ids(rbk, world) == rank).all() and (_partition_ids(rpk, world) == rank).all()   # co-partitioned
    # inside every source block the original order is preserved
    start = 0
    for src, n in enumerate(rc):
        blk = rpp[start:start + n] - src * 10**9
        assert (np.diff(blk) > 0).all()
        start += n
    bi, pi, _, _ = O.hash_join_indices([(rbk, None)], [(rpk, None)])
    local = np.stack([rbk[bi], rbp[bi], rpp[pi]], axis=1) if len(bi) else np.zeros((0, 3), np.int64)
    out_q.put((rank, bk, bp, pk, pp, local))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_then_local_join_equals_global_join():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle as O
    bk = np.concatenate([r[1] for r in res]); bp = np.concatenate([r[2] for r in res])
    pk = np.concatenate([r[3] for r in res]); pp = np.concatenate([r[4] for r in res])
    bi, pi, _, _ = O.hash_join_indices([(bk, None)], [(pk, None)])
    glob = np.stack([bk[bi], bp[bi], pp[pi]], axis=1)
    loc = np.concatenate([r[5] for r in res])
    assert len(glob) == len(loc) > 0
    assert np.array_equal(glob[np.lexsort(glob.T[::-1])], loc[np.lexsort(loc.T[::-1])])


def test_plan_all_to_all_offsets():
    from datafusion_b200 import exchange
    sc, rc, so, ro = exchange.plan_all_to_all([3, 0, 5], [1, 2, 0])
    assert so.tolist() == [0, 3, 3, 8] and ro.tolist() == [0, 1, 3, 3]


def test_peer_chunk_layout_is_a_disjoint_cover():
    """layout of the chunked peer exchange: every (src, chunk, dst) block lands in its own range of the receiver's buffer,
    the ranges tile [0, total) exactly, and chunk c of all sources is one contiguous slice (what the pipelined join consumes)"""
    from datafusion_b200.exchange import peer_chunk_layout
    rng = np.random.default_rng(3)
    for world, chunks in ((2, 1), (2, 4), (8, 3), (5, 7)):
        m = rng.integers(0, 50, size=(world, chunks, world))
        m[rng.integers(0, world), rng.integers(0, chunks), :] = 0        # an empty block
        layouts = [peer_chunk_layout(m, r) for r in range(world)]
        for dst in range(world):
            total = int(m[:, :, dst].sum())
            owner = np.full(total, -1)
            for src in range(world):
                dst_row = layouts[src][0]
                for c in range(chunks):
                    lo = int(dst_row[c][dst]); hi = lo + int(m[src, c, dst])
                    assert (owner[lo:hi] == -1).all()
                    owner[lo:hi] = c * world + src
            assert (owner >= 0).all()
            assert (np.diff(owner) >= 0).all()                            # chunk-major, then source rank
            _, start, rows, mx = layouts[dst]
            for c in range(chunks):
                assert rows[c] == m[:, c, dst].sum()
                assert start[c] == m[:, :c, dst].sum()
            assert mx == m.sum(axis=(0, 1)).max()


def _q3_worker(rank, world, port, out_q):
    """the partitioned multi-GPU Q3 plan (scripts/q3_multi_gpu.py) with the oracle's CPU operators in place of the kernels: same sharding,
    same exchanges (customer keys all-gathered; qualified orders and filtered lineitems hash-exchanged to their owners; groups owned by
    the owner of their order key), same fingerprint reduction"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datafusion_b200 import exchange
    from oracle import oracle as O
    sf = 0.02
    t = O.q3_generate(sf * world, threads=1)                                   # the global database; this rank owns a contiguous row range of each table
    sl = lambda a, n: a[rank * n:(rank + 1) * n]
    nc, no, nl = len(t["c_custkey"]) // world, len(t["o_orderkey"]) // world, len(t["l_orderkey"]) // world
    dev = torch.device("cpu")

    def a2a(cols, key):
        pid = _partition_ids(cols[key], world)
        order = np.argsort(pid, kind="stable")
        sc_ = np.bincount(pid, minlength=world).tolist()
        rc_ = exchange.exchange_counts(dist, sc_, dev)
        sc, rc, so, ro = exchange.plan_all_to_all(sc_, rc_)
        got = exchange.all_to_all_columns(dist, [torch.from_numpy(np.ascontiguousarray(c[order])) for c in cols], sc, rc)
        return [g.numpy() for g in got]
    # customer: filter -> keys -> all-gather (CollectLeft)
    ck = sl(t["c_custkey"], nc)[sl(t["c_mktsegment"], nc) == 1]
    cnt = torch.tensor([len(ck)]); allc = [torch.zeros_like(cnt) for _ in range(world)]; dist.all_gather(allc, cnt)
    mx = max(int(c.item()) for c in allc)
    pad = torch.zeros(mx, dtype=torch.int64); pad[:len(ck)] = torch.from_numpy(ck)
    allk = [torch.zeros_like(pad) for _ in range(world)]; dist.all_gather(allk, pad)
    cust = np.concatenate([k.numpy()[:int(c.item())] for k, c in zip(allk, allc)])
    # orders: filter + semi -> exchange by o_orderkey
    ok, oc, od, op_ = sl(t["o_orderkey"], no), sl(t["o_custkey"], no), sl(t["o_orderdate"], no), sl(t["o_shippriority"], no)
    m = (od < O.Q3_CUT) & np.isin(oc, cust)
    qo = a2a([ok[m], od[m].astype(np.int64), op_[m].astype(np.int64)], 0)
    # lineitem: filter -> exchange by l_orderkey (the membership filter only removes rows without a partner: same result)
    lk, lp, ld, ls = sl(t["l_orderkey"], nl), sl(t["l_extendedprice"], nl), sl(t["l_discount"], nl), sl(t["l_shipdate"], nl)
    m = ls > O.Q3_CUT
    ql = a2a([lk[m], lp[m], ld[m]], 0)
    # owner: Inner join + group-by (the oracle's operators)
    j = O.hash_join([(qo[0], None), (qo[1], None), (qo[2], None)], [(ql[0], None), (ql[1], None), (ql[2], None)], [0], [0], [1, 0, 0, 1, 1], [0, 1, 2, 1, 2])
    rev = (j[3][0].astype(np.uint64) * (100 - j[4][0]).astype(np.uint64)).view(np.int64)
    keys, res = O.group_by([j[0], j[1], j[2]], [(O.A_SUM, (rev, None), None)]) if len(rev) else ([(np.zeros(0, np.int64), None)] * 3, [{"i": np.zeros(0, np.int64)}])
    usum = lambda a: int(np.asarray(a).astype(np.int64).view(np.uint64).sum(dtype=np.uint64))
    fp_local = [len(keys[0][0]), usum(keys[0][0]), usum(keys[1][0]), usum(keys[2][0]), usum(res[0]["i"]), len(rev), len(qo[0])]
    fp = exchange.allgather_wrapping_sum(dist, fp_local, dev)
    if rank == 0:
        efp, ejoined, eorders = O.q3_stream_fingerprint(sf * world, threads=2)
        out_q.put((fp, efp + [ejoined, eorders]))
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_q3_plan_world_size_2_matches_the_global_evaluation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_q3_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, exp = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == exp and got[0] > 100
