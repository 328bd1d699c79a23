"""The ctypes binding's marshalling helpers (datafusion_b200/capi.py) — what crosses the C ABI must have exactly Arrow's layout
(validity = LSB-numbered bitmap, Decimal128 = 16-byte two's complement little endian, BOOL bit-packed) and the header's encodings
(DFGPU_DECIMAL128_TYPE, a 128-bit literal split over lit_i64 / lit_f64).  No GPU needed; pyarrow is the independent reference."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from datafusion_b200 import capi as D


def test_pack_bits_is_arrows_validity_bitmap_layout():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000):
        m = rng.random(n) > 0.4
        b = D.pack_bits(m)
        assert len(b) % 8 == 0 and len(b) >= 8                      # padded for the library's 64-bit bitmap loads
        assert np.array_equal(D.unpack_bits(b, n), m)
        if n:
            ref = np.frombuffer(pa.array([1 if x else None for x in m.tolist()], pa.int8()).buffers()[0] or b"", np.uint8) if not m.all() else None
            if ref is not None:
                assert np.array_equal(b[:len(ref)][: (n + 7) // 8], ref[: (n + 7) // 8])
    assert np.array_equal(D.unpack_bits(D.pack_bits(np.array([0, 1, 1, 0, 1], bool)), 3, offset=1), [True, True, False])


def test_decimal_words_are_arrows_decimal128_buffer():
    vals = [0, 1, -1, 12345678901234567890123, -12345678901234567890123, (1 << 127) - 1, -(1 << 127), 10**37, -(10**37)]
    w = D.decimal_to_words(vals)
    assert w.dtype == np.uint64 and w.shape == (len(vals), 2)
    assert D.words_to_decimal(w) == vals
    import decimal
    with decimal.localcontext() as dctx:
        dctx.prec = 60
        arr = pa.array([decimal.Decimal(v) for v in vals[:5] + vals[7:]], pa.decimal128(38, 0))
    ref = np.frombuffer(arr.buffers()[1], np.uint64).reshape(-1, 2)
    assert np.array_equal(D.decimal_to_words(vals[:5] + vals[7:]), ref)


def test_host_column_describes_arrow_buffers():
    v = np.arange(10, dtype=np.int32); valid = np.array([1, 1, 0, 1, 1, 1, 0, 1, 1, 1], bool)
    h = D.HostColumn(v, valid)
    c = h.c()
    assert (c.type, c.length, c.offset, c.null_count) == (D.INT32, 10, 0, 2) and c.values and c.validity
    assert np.array_equal(np.ctypeslib.as_array(C.cast(c.values, C.POINTER(C.c_int32)), (10,)), v)
    assert D.HostColumn(v).c().validity is None and D.HostColumn(v).null_count == 0
    b = D.HostColumn(np.array([True, False, True, True]))
    assert b.type == D.BOOL and b._values[0] == 0b1101                         # BOOL values are bit-packed, LSB first
    d = D.HostColumn([5, -7], None, D.decimal128(15, 2))
    assert d._values.shape == (2, 2) and D.words_to_decimal(d._values) == [5, -7]
    t = D.HostColumn(np.array([1, 2], np.int32), None, D.DATE32)
    assert t.type == D.DATE32 and t._values.dtype == np.int32


@pytest.mark.parametrize("value", [0, 1, -1, 123456789012345678901234567890, -(10**30), (1 << 127) - 1, -(1 << 127)])
def test_decimal_literal_is_split_over_the_two_literal_fields(value):
    nodes = D.expr_nodes([(D.EXPR_LITERAL, 0, D.decimal128(38, 4), 0, value, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, -5, 0.0), (D.EXPR_LITERAL, 0, D.FLOAT64, 0, 0, 2.5)])
    raw = bytes(C.string_at(C.addressof(nodes[0]) + D.ExprNode.lit_i64.offset, 16))
    assert int.from_bytes(raw, "little", signed=True) == value             # lit_i64 | lit_f64 = the 16 little-endian bytes of the i128
    assert (nodes[0].kind, nodes[0].type, nodes[0].is_null) == (D.EXPR_LITERAL, D.decimal128(38, 4), 0)
    assert nodes[1].lit_i64 == -5 and nodes[2].lit_f64 == 2.5


def test_type_tables_are_consistent():
    for tid, npt in D.NP_OF_TYPE.items():
        if tid == D.BOOL or D.type_base(tid) == D.DECIMAL128:
            continue
        assert D.WIDTH[tid] == np.dtype(npt).itemsize, tid
    assert D.WIDTH[D.decimal128(15, 2)] == 16 and D.WIDTH[D.decimal128(38, -3)] == 16
    assert D.decimal_precision_scale(D.decimal128(38, -3)) == (38, -3) and D.type_base(D.decimal128(9, 9)) == D.DECIMAL128
    for npt, tid in D.TYPE_OF_NP.items():
        if tid != D.BOOL:                                                   # BOOL is bit-packed: no numpy element type
            assert np.dtype(D.NP_OF_TYPE[tid]) == np.dtype(npt)
