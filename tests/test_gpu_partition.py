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


@pytest.mark.parametrize("n_parts", [2, 8, 5])
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
