"""Join keys that do not fit the exact 64-bit tag (several columns wider than 64 bits together, Decimal128 keys, more than four key
columns): the table is keyed by a hash of the key columns and key equality is checked on the candidate pairs — the reference's own
scheme (lookup by hash, then equal_rows_arr, joins/utils.rs:2191-2257).  Results must equal the oracle's multi-column join."""
import numpy as np
import pytest

from datafusion_b200 import capi as D
from oracle import oracle as O
from harness import assert_cols_equal, gpu_hash_join
from test_oracle_golden import out_mapping

pytestmark = pytest.mark.gpu
JT = {"Inner": O.J_INNER, "Left": O.J_LEFT, "Right": O.J_RIGHT, "Full": O.J_FULL, "LeftSemi": O.J_LEFT_SEMI, "RightSemi": O.J_RIGHT_SEMI,
      "LeftAnti": O.J_LEFT_ANTI, "RightAnti": O.J_RIGHT_ANTI, "LeftMark": O.J_LEFT_MARK, "RightMark": O.J_RIGHT_MARK}
GJT = {"Inner": D.JOIN_INNER, "Left": D.JOIN_LEFT, "Right": D.JOIN_RIGHT, "Full": D.JOIN_FULL, "LeftSemi": D.JOIN_LEFT_SEMI,
       "RightSemi": D.JOIN_RIGHT_SEMI, "LeftAnti": D.JOIN_LEFT_ANTI, "RightAnti": D.JOIN_RIGHT_ANTI, "LeftMark": D.JOIN_LEFT_MARK, "RightMark": D.JOIN_RIGHT_MARK}
ORDERED = ("Inner", "RightSemi", "RightAnti", "RightMark")


def wide_tables(rng, nb, npr, n_distinct, dup, null_frac):
    """keys (a: int64, b: int64, c: int32): 160 bits together; (a, b) alone does not identify the key when c differs"""
    ka = rng.integers(-2**62, 2**62, n_distinct).astype(np.int64)
    kb = rng.integers(0, 3, n_distinct).astype(np.int64) * (2**40)
    kc = rng.integers(0, 4, n_distinct).astype(np.int32)
    bi = np.resize(np.repeat(np.arange(n_distinct // 2), dup), nb); rng.shuffle(bi)       # the build side holds half of the key universe
    pi = rng.integers(0, n_distinct, npr)
    nul = lambda n: None if null_frac == 0 else rng.random(n) >= null_frac
    build = [(ka[bi], nul(nb)), (kb[bi], nul(nb)), (kc[bi], None), (np.arange(nb, dtype=np.int64), None)]
    probe = [(ka[pi], nul(npr)), (kb[pi], None), ((kc[pi] + (rng.random(npr) < 0.2)).astype(np.int32), nul(npr)), (np.arange(npr, dtype=np.int64) * 3, None)]
    return build, probe


@pytest.mark.parametrize("jt", list(JT))
@pytest.mark.parametrize("dup,null_frac", [(1, 0.0), (3, 0.08)])
def test_wide_three_column_key_vs_oracle(gpu_ctx, jt, dup, null_frac):
    rng = np.random.default_rng(hash((jt, dup)) % 2**32)
    build, probe = wide_tables(rng, 3000, 11000, 2400, dup, null_frac)
    side, idx = out_mapping(jt, 4, 4)
    exp = O.hash_join(build, probe, [0, 1, 2], [0, 1, 2], side, idx, join_type=JT[jt], phj_threshold=0, phj_density=float("inf"))
    got = gpu_hash_join(gpu_ctx, build, probe, [0, 1, 2], [0, 1, 2], side, idx, GJT[jt], probe_batch_rows=4000, build_batch_rows=1100)
    assert_cols_equal(got, exp, ordered=jt in ORDERED, what=f"{jt} dup={dup} nulls={null_frac}")


def test_wide_two_int64_keys_device_path_and_metrics(gpu_ctx):
    rng = np.random.default_rng(3)
    build, probe = wide_tables(rng, 5000, 20000, 4000, 2, 0.0)
    side, idx = out_mapping("Inner", 4, 4)
    exp = O.hash_join(build, probe, [0, 1], [0, 1], side, idx)
    got, h = gpu_hash_join(gpu_ctx, build, probe, [0, 1], [0, 1], side, idx, device=True, probe_batch_rows=6000, return_handle=True)
    assert_cols_equal(got, exp, ordered=True)
    assert h.metric("array_map_created_count") == 0 and h.metric("output_rows") == len(exp[0][0]) > 1000
    h.close()


@pytest.mark.parametrize("jt", ["Inner", "Left", "RightAnti", "Full"])
def test_wide_keys_null_equals_null(gpu_ctx, jt):
    rng = np.random.default_rng(4)
    build, probe = wide_tables(rng, 600, 2500, 300, 2, 0.15)
    side, idx = out_mapping(jt, 4, 4)
    exp = O.hash_join(build, probe, [0, 1, 2], [0, 1, 2], side, idx, join_type=JT[jt], null_equals_null=True, phj_threshold=0, phj_density=float("inf"))
    got = gpu_hash_join(gpu_ctx, build, probe, [0, 1, 2], [0, 1, 2], side, idx, GJT[jt], D.NULL_EQUALS_NULL)
    assert_cols_equal(got, exp, ordered=jt in ORDERED, what=jt)


@pytest.mark.parametrize("jt", ["Inner", "Left", "RightSemi", "LeftAnti", "Full"])
def test_wide_keys_with_a_join_filter(gpu_ctx, jt):
    """a user JoinFilter on top of the key-equality conjunct: build.payload % 5 > probe.payload % 3"""
    rng = np.random.default_rng(8)
    build, probe = wide_tables(rng, 2500, 9000, 900, 3, 0.05)
    side, idx = out_mapping(jt, 4, 4)
    onodes = [(O.E_COLUMN, 0, None, 0, 0), (O.E_LITERAL, 0, np.int64, 0, 5), (O.E_BINARY, O.OP_MODULO, None, 0, 0),
              (O.E_COLUMN, 1, None, 0, 0), (O.E_LITERAL, 0, np.int64, 0, 3), (O.E_BINARY, O.OP_MODULO, None, 0, 0), (O.E_BINARY, O.OP_GT, None, 0, 0)]
    gnodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, 5, 0.0), (D.EXPR_BINARY, D.OP_MODULO, 0, 0, 0, 0.0),
              (D.EXPR_COLUMN, 1, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, 3, 0.0), (D.EXPR_BINARY, D.OP_MODULO, 0, 0, 0, 0.0), (D.EXPR_BINARY, D.OP_GT, 0, 0, 0, 0.0)]
    exp = O.hash_join(build, probe, [0, 1, 2], [0, 1, 2], side, idx, join_type=JT[jt], filter=([0, 1], [3, 3], onodes), phj_threshold=0, phj_density=float("inf"))
    got = gpu_hash_join(gpu_ctx, build, probe, [0, 1, 2], [0, 1, 2], side, idx, GJT[jt], probe_batch_rows=4000, filter=([0, 1], [3, 3], gnodes))
    assert_cols_equal(got, exp, ordered=jt in ORDERED, what=jt)


@pytest.mark.parametrize("jt", ["Inner", "Right", "LeftSemi", "RightAnti"])
def test_decimal128_join_key(gpu_ctx, jt):
    """one Decimal128(38,4) key column: 128 bits; the oracle joins on the (low word, high word) pair"""
    import random
    r = random.Random(9)
    universe = [r.randint(-10**37, 10**37) for _ in range(500)] + [-1, 0, 1, (1 << 64), -(1 << 64), (1 << 64) + 1]
    nb, npr = 700, 4000
    bvals = [universe[r.randrange(len(universe) // 2)] for _ in range(nb)]
    pvals = [universe[r.randrange(len(universe))] for _ in range(npr)]
    bvalid = np.array([r.random() > 0.05 for _ in range(nb)], bool); pvalid = np.array([r.random() > 0.05 for _ in range(npr)], bool)
    bw, pw = D.decimal_to_words(bvals), D.decimal_to_words(pvals)
    bpay, ppay = np.arange(nb, dtype=np.int64), np.arange(npr, dtype=np.int64) * 7
    t = D.decimal128(38, 4)
    side, idx = out_mapping(jt, 1, 1)          # payload columns only
    omap = lambda s, i: (s, [1 if x == 0 else x for x in i])
    # oracle: keys = (lo, hi) as int64 columns, payload = column 2
    ob = [(bw[:, 0].view(np.int64).copy(), bvalid), (bw[:, 1].view(np.int64).copy(), bvalid), (bpay, None)]
    op = [(pw[:, 0].view(np.int64).copy(), pvalid), (pw[:, 1].view(np.int64).copy(), pvalid), (ppay, None)]
    oidx = [2 if s in (0, 1) else 0 for s in side]
    exp = O.hash_join(ob, op, [0, 1], [0, 1], side, oidx, join_type=JT[jt], phj_threshold=0, phj_density=float("inf"))
    gidx = [1 if s in (0, 1) else 0 for s in side]
    got = gpu_hash_join(gpu_ctx, [(bw, bvalid), (bpay, None)], [(pw, pvalid), (ppay, None)], [0], [0], side, gidx, GJT[jt],
                        build_types=[t, D.INT64], probe_types=[t, D.INT64], probe_batch_rows=1500)
    assert_cols_equal(got, exp, ordered=jt in ORDERED, what=jt)
    assert len(exp[0][0]) > 100


def test_five_narrow_key_columns(gpu_ctx):
    """more key columns than the packed tag supports (4), although they are narrow: the wide-key path by column count"""
    rng = np.random.default_rng(12)
    nb, npr = 2000, 8000
    bk = [rng.integers(0, 4, nb).astype(np.int8) for _ in range(5)]
    pk = [rng.integers(0, 4, npr).astype(np.int8) for _ in range(5)]
    build = [(k, None) for k in bk] + [(np.arange(nb, dtype=np.int64), None)]
    probe = [(k, None) for k in pk] + [(np.arange(npr, dtype=np.int64), None)]
    side, idx = [0, 1], [5, 5]
    exp = O.hash_join(build, probe, [0, 1, 2, 3, 4], [0, 1, 2, 3, 4], side, idx)
    got = gpu_hash_join(gpu_ctx, build, probe, [0, 1, 2, 3, 4], [0, 1, 2, 3, 4], side, idx, probe_batch_rows=3000)
    assert_cols_equal(got, exp, ordered=True)
    assert len(exp[0][0]) > 5000
