"""The paired-accumulator group-by kernel (agg_update_pair_kernel; DFGPU_AGG_PAIRED = 1 (default) | 2 | 3 | 4, 0 = the kernel with one RED
per aggregate and row): the two-aggregate fast path
(SUM + COUNT over a non-null int64 column, the C3 shape; Final-mode merges of two states) with one L2 reduction request
per row.  Same outputs as the default kernel and as the oracle: bit-exact, rows compared sorted (aggregation_fuzzer/mod.rs:59-86).
The mode is read when the handle is created, so the tests switch it per handle."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import assert_cols_equal, gpu_group_by

pytestmark = pytest.mark.gpu


def oracle_sum_count(g, v, vv=None):
    keys, res = O.group_by([(g, None)], [(O.A_SUM, (v, vv), None), (O.A_COUNT, (v, vv), None)])
    return [keys[0]] + O.agg_output_columns(O.A_SUM, res[0], np.int64, False) + O.agg_output_columns(O.A_COUNT, res[1], np.int64, False)


@pytest.mark.parametrize("mode", ["0", "1", "2", "3", "4"])
@pytest.mark.parametrize("n,groups,batch_rows,hint", [(100_003, 700, None, 0), (1_000_001, 300_000, 250_000, 0), (2_000_000, 50_000, 999_983, 50_000), (37, 5, 7, 0)])
def test_paired_sum_count_equals_oracle(gpu_ctx, monkeypatch, mode, n, groups, batch_rows, hint):
    monkeypatch.setenv("DFGPU_AGG_PAIRED", mode)
    rng = np.random.default_rng(n + groups)
    g = (rng.integers(0, groups, n).astype(np.int64) * 1_000_003) - 17        # sparse, some negative
    g[rng.integers(0, n, 3)] = -1                                              # all-ones key: the table's special slot
    v = rng.integers(-2**62, 2**62, n).astype(np.int64)                        # sums wrap (sum.rs:316)
    got = gpu_group_by(gpu_ctx, [(g, None), (v, None)], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], batch_rows=batch_rows, device=True, capacity_hint=hint)
    assert_cols_equal(got, oracle_sum_count(g, v), ordered=False, what=f"paired mode {mode}")


@pytest.mark.parametrize("mode", ["1", "2", "3", "4"])
def test_paired_then_generic_batches_fold_correctly(gpu_ctx, monkeypatch, mode):
    """batches without NULLs take the paired kernel, a batch with NULLs takes the generic kernel on the per-aggregate arrays,
    then paired again: the deltas must be folded at every switch, at table growth and before the emit"""
    monkeypatch.setenv("DFGPU_AGG_PAIRED", mode)
    rng = np.random.default_rng(5)
    n = 600_000
    g = rng.integers(0, 120_000, n).astype(np.int64) * 7
    v = rng.integers(-10**12, 10**12, n).astype(np.int64)
    vv = np.ones(n, bool); vv[200_000:400_000] = rng.random(200_000) > 0.3   # the middle batch has NULLs
    a = D.AggHandle(gpu_ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], D.AGG_SINGLE, 8192, 0)
    keep = []
    for s, e in ((0, 200_000), (200_000, 400_000), (400_000, 600_000)):
        has_null = not vv[s:e].all()
        hc = [D.HostColumn(g[s:e]), D.HostColumn(v[s:e], vv[s:e] if has_null else None)]
        dc = [D.DeviceColumn.from_host(gpu_ctx, h) for h in hc]; keep.append(dc)
        a.push_device(dc)
    a.finish()
    from harness import batches_to_cols
    outs = a.drain(host=False)
    got = batches_to_cols(outs, outs[0].num_columns)
    a.close()
    assert_cols_equal(got, oracle_sum_count(g, v, vv), ordered=False, what="paired / generic / paired")


@pytest.mark.parametrize("mode", ["0", "1", "3"])
def test_paired_partial_then_final_equals_single(gpu_ctx, monkeypatch, mode):
    monkeypatch.setenv("DFGPU_AGG_PAIRED", mode)
    rng = np.random.default_rng(11)
    n = 400_000
    g = rng.integers(0, 9_000, n).astype(np.int64)
    v = rng.integers(-2**40, 2**40, n).astype(np.int64)
    aggs = [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)]
    parts = []
    for s, e in ((0, 150_000), (150_000, 400_000)):
        parts.append(gpu_group_by(gpu_ctx, [(g[s:e], None), (v[s:e], None)], [0], aggs, mode=D.AGG_PARTIAL, device=True))
    assert all(p[i][1] is None for p in parts for i in range(3))               # no NULL inputs: no NULL states
    st = [(np.concatenate([p[i][0] for p in parts]), None) for i in range(3)]   # [key, sum state, count state]: Final merges both with `+=`
    fin = gpu_group_by(gpu_ctx, st, [0], [(D.AGG_SUM, -1, -1), (D.AGG_COUNT, -1, -1)], mode=D.AGG_FINAL, device=True)
    assert_cols_equal(fin, oracle_sum_count(g, v), ordered=False, what="partial -> final, paired merge")
