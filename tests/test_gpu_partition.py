"""Hash partitioning ahead of the exchange (RepartitionExec::BatchPartitioner::Hash, repartition/mod.rs:1097-1145):
every row lands in partition hash(key) % n, rows keep their input order inside a partition (stable), payload
columns travel with their keys, and equal keys always share a partition (co-partitioning of both join sides)."""
import numpy as np
import pytest

from datafusion_b200 import capi as D

pytestmark = pytest.mark.gpu
SEED_EXCHANGE = np.uint64(0x9E3779B97F4A7C15)


def mix64(x):
    x = x.copy()
    x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


@pytest.mark.parametrize("n_parts", [1, 2, 8, 5, 13, 32])
def test_hash_partition_is_stable_and_complete(gpu_ctx, n_parts):
    rng = np.random.default_rng(n_parts)
    n = 100_007
    k = rng.integers(-2**62, 2**62, n).astype(np.int64); v = np.arange(n, dtype=np.int64); w = rng.integers(0, 100, n).astype(np.int32)
    cols = [D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(x)) for x in (k, v, w)]
    batch, offs = D.hash_partition_device(gpu_ctx, cols, [0], n_parts)
    gk, gv, gw = (batch.column_numpy(i)[0] for i in range(3))
    with np.errstate(over="ignore"):
        h = mix64(k.view(np.uint64) + SEED_EXCHANGE)
        # partition = fastrange(hash, n) = floor(hash * n / 2^64)  (the reference uses hash % n, repartition/mod.rs:875-935;
        # which partition a row lands in is not observable in operator output, only co-partitioning is)
        pid = np.array([(int(x) * n_parts) >> 64 for x in h], dtype=np.int64)
    assert offs[0] == 0 and offs[-1] == n
    order = np.argsort(pid, kind="stable")
    assert np.array_equal(gk, k[order]) and np.array_equal(gv, v[order]) and np.array_equal(gw, w[order])
    assert [offs[p + 1] - offs[p] for p in range(n_parts)] == np.bincount(pid, minlength=n_parts).tolist()


def test_partitioned_join_equals_global_join(gpu_ctx):
    # PartitionMode::Partitioned (hash_join/exec.rs:1312-1325): join partition by partition == the global join
    from oracle import oracle as O
    from harness import assert_cols_equal, gpu_hash_join
    rng = np.random.default_rng(2)
    bk = rng.permutation(60000)[:20000].astype(np.int64) * 31; bp = rng.integers(0, 1 << 40, 20000).astype(np.int64)
    pk = rng.integers(0, 60000, 150000).astype(np.int64) * 31; pp = np.arange(150000, dtype=np.int64)
    P = 4
    parts = []
    for cols in ([bk, bp], [pk, pp]):
        dc = [D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(x)) for x in cols]
        batch, offs = D.hash_partition_device(gpu_ctx, dc, [0], P)
        parts.append(([batch.column_numpy(i)[0] for i in range(2)], offs))
    outs = []
    for p in range(P):
        b = [(parts[0][0][c][parts[0][1][p]:parts[0][1][p + 1]], None) for c in range(2)]
        q = [(parts[1][0][c][parts[1][1][p]:parts[1][1][p + 1]], None) for c in range(2)]
        outs.append(gpu_hash_join(gpu_ctx, b, q, [0], [0], [0, 0, 1, 1], [0, 1, 0, 1]))
    got = [(np.concatenate([o[c][0] for o in outs]), None) for c in range(4)]
    exp = O.hash_join([(bk, None), (bp, None)], [(pk, None), (pp, None)], [0], [0], [0, 0, 1, 1], [0, 1, 0, 1])
    assert_cols_equal(got, exp, ordered=False)


@pytest.mark.parametrize("n_parts", [3, 8, 16])
def test_hash_partition_int32_key_many_widths(gpu_ctx, n_parts):
    # non-8-byte key (generic hash path) with 1/2/4/8-byte payload columns: same stable-order contract
    rng = np.random.default_rng(40 + n_parts)
    n = 70_001
    k = rng.integers(0, 5000, n).astype(np.int32)
    cols_np = [k, rng.integers(-100, 100, n).astype(np.int8), rng.integers(0, 60000, n).astype(np.uint16), rng.random(n).astype(np.float32),
               np.arange(n, dtype=np.int64), rng.random(n)]
    cols = [D.DeviceColumn.from_host(gpu_ctx, D.HostColumn(x)) for x in cols_np]
    batch, offs = D.hash_partition_device(gpu_ctx, cols, [0], n_parts)
    with np.errstate(over="ignore"):
        h = mix64(k.astype(np.uint64) + SEED_EXCHANGE)
    pid = np.array([(int(x) * n_parts) >> 64 for x in h], dtype=np.int64)
    order = np.argsort(pid, kind="stable")
    for i, x in enumerate(cols_np):
        assert np.array_equal(batch.column_numpy(i)[0], x[order]), f"column {i}"
    assert offs[-1] == n and [offs[p + 1] - offs[p] for p in range(n_parts)] == np.bincount(pid, minlength=n_parts).tolist()


@pytest.mark.parametrize("n_parts,n_chunks", [(2, 1), (8, 4), (5, 7), (12, 3)])
def test_chunked_peer_plan_scatters_every_chunk_to_its_block(gpu_ctx, n_parts, n_chunks):
    """dfgpu_partition_plan_create_chunked / _scatter_peer_chunk with every "peer" buffer on this GPU: chunk c of
    partition p must land, in input order, at the offset the caller passed — the contract PartitionedHashJoin's
    receive layout (exchange.peer_chunk_layout) is built on."""
    import ctypes as C
    ctx = gpu_ctx
    rng = np.random.default_rng(n_parts * 10 + n_chunks)
    n = 50_000 + n_parts
    k = rng.integers(-2**62, 2**62, n).astype(np.int64); v = np.arange(n, dtype=np.int64)
    cols = [D.DeviceColumn.from_host(ctx, D.HostColumn(x)) for x in (k, v)]
    counts = (C.c_int64 * (n_parts * n_chunks))()
    plan = C.c_void_p()
    ctx.check(ctx.lib.dfgpu_partition_plan_create_chunked(ctx.h, D._cols(cols), 2, D._i32arr([0]), 1, n_parts, n_chunks, counts, C.byref(plan)))
    try:
        cnt = np.array(list(counts), dtype=np.int64).reshape(n_chunks, n_parts)
        with np.errstate(over="ignore"):
            h = mix64(k.view(np.uint64) + SEED_EXCHANGE)
        pid = np.array([(int(x) * n_parts) >> 64 for x in h], dtype=np.int64)
        tile = 2048
        ntiles = (n + tile - 1) // tile
        bounds = [min(n, (ntiles * c // n_chunks) * tile) for c in range(n_chunks)] + [n]
        for c in range(n_chunks):
            assert cnt[c].tolist() == np.bincount(pid[bounds[c]:bounds[c + 1]], minlength=n_parts).tolist()
        # one receive buffer per partition, chunk-major layout with a 3-row gap before every block
        gap = 3
        bufs = [[D.DeviceBuffer(ctx, (int(cnt[:, p].sum()) + gap * n_chunks + 1) * 8) for _ in range(2)] for p in range(n_parts)]
        bases = (C.c_void_p * (n_parts * 2))(*[bufs[p][c].ptr for p in range(n_parts) for c in range(2)])
        starts = np.zeros((n_chunks, n_parts), dtype=np.int64)
        for p in range(n_parts):
            run = 0
            for c in range(n_chunks):
                run += gap
                starts[c, p] = run
                run += cnt[c, p]
        for c in reversed(range(n_chunks)):      # any order of chunk calls must work
            rows = (C.c_int64 * n_parts)(*[int(x) for x in starts[c]])
            ctx.check(ctx.lib.dfgpu_partition_plan_scatter_peer_chunk(plan, c, bases, rows))
        ctx.sync()
        for p in range(n_parts):
            for c in range(n_chunks):
                sel = np.nonzero(pid[bounds[c]:bounds[c + 1]] == p)[0] + bounds[c]
                for ci, src in enumerate((k, v)):
                    got = ctx.to_host(bufs[p][ci].ptr + int(starts[c, p]) * 8, len(sel) * 8).view(np.int64)
                    assert np.array_equal(got, src[sel]), (p, c, ci)
    finally:
        ctx.lib.dfgpu_partition_plan_destroy(plan)
