// batch.cuh — record batches (device / pinned host), transfers, bitmap utilities, take/concat.
#pragma once
#include "common.cuh"
#include <deque>

namespace dfgpu {

struct HCol {
  int type = 0;
  int64_t length = 0;
  int64_t null_count = -1;
  std::shared_ptr<HostBuf> values, validity;
};

}  // namespace dfgpu

struct dfgpu_batch {
  dfgpu_ctx* ctx = nullptr;
  int64_t rows = 0;
  bool host = false;
  std::vector<dfgpu::DCol> cols;  // device columns
  std::vector<dfgpu::HCol> hcols;  // host (pinned) columns
};

namespace dfgpu {

using BatchPtr = std::unique_ptr<dfgpu_batch>;

// ---- views over caller-provided columns --------------------------------------------------
// DCol convention: `values` points at logical element 0 for fixed-width types; for BOOL the value
// bit i lives at bit (offset + i) of `values`; validity bit i lives at bit (offset + i).
inline DCol device_view(const dfgpu_column& c) {
  DCol d;
  d.type = c.type; d.length = c.length; d.null_count = c.null_count;
  d.validity = (c.null_count == 0) ? nullptr : c.validity;
  if (c.type == DFGPU_BOOL) {
    d.values = c.values; d.offset = c.offset;
  } else {
    int w = type_width(c.type);
    DF_CHECK(w > 0, DFGPU_ERR_UNSUPPORTED, "unsupported column type");
    d.values = (const char*)c.values + c.offset * w;
    d.offset = c.offset;  // applies to validity only
  }
  if (!d.validity) d.offset = (c.type == DFGPU_BOOL) ? c.offset : 0;
  return d;
}

// NOTE on DCol.offset for non-BOOL: it is the validity bit offset.  For BOOL columns it is the bit
// offset of both the values and the validity bitmap (Arrow slices share one logical offset).

DCol upload_column(dfgpu_ctx* ctx, const dfgpu_column& hc);                 // H2D (async on ctx stream)
DCol copy_column_device(dfgpu_ctx* ctx, const DCol& src);                     // D2D owned copy, offset normalised to <8 bits
BatchPtr to_host_batch(dfgpu_ctx* ctx, const dfgpu_batch& dev);               // D2H into pinned buffers (syncs)
DCol concat_columns(dfgpu_ctx* ctx, const std::vector<DCol>& parts, int type);  // arrow concat_batches
DCol slice_column(const DCol& c, int64_t start, int64_t len);

// out[i] = src[idx[i]]; idx == 0xFFFFFFFF yields NULL (outer-join padding, joins/utils.rs:1332-1387).
// idx_may_be_null: whether any idx can be the null marker.
DCol take_column(dfgpu_ctx* ctx, const DCol& src, const uint32_t* idx_dev, int64_t n, bool idx_may_be_null);
// all-null column of n rows (new_null_array)
DCol null_column(dfgpu_ctx* ctx, int type, int64_t n);

// popcount of a device bitmap range
int64_t count_set_bits(dfgpu_ctx* ctx, const uint8_t* bm, int64_t bit_offset, int64_t n);

// dst bits [dst_off, dst_off+n) |= src bits [src_off, src_off+n)   (dst must be pre-zeroed there)
void bitmap_or_copy(dfgpu_ctx* ctx, uint8_t* dst, int64_t dst_off, const uint8_t* src, int64_t src_off, int64_t n);
void bitmap_set_range(dfgpu_ctx* ctx, uint8_t* dst, int64_t dst_off, int64_t n);

// stream compaction of row indices by a flag bitmap: out_idx = ascending positions of set (want_set=1)
// or clear (want_set=0) bits among the first n bits of `words`.  Returns the count.
int64_t compact_flag_indices(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, DevBuf* out_idx);
int64_t compact_count(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, DevBuf* tiles);
void compact_emit(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, const DevBuf& tiles, uint32_t* dst);
void fill_iota(dfgpu_ctx* ctx, uint32_t* out, int64_t n);

// ---- device bit loads (safe at buffer edges: only touches bytes that hold requested bits) ----
__device__ __forceinline__ uint32_t load_bits32(const uint8_t* bm, int64_t bit, int nbits) {
  // returns bits [bit, bit+nbits) in the low nbits of the result, nbits in [1,32]
  int64_t byte0 = bit >> 3;
  int sh = (int)(bit & 7);
  int nbytes = (sh + nbits + 7) >> 3;  // <= 5
  uint64_t acc = 0;
#pragma unroll
  for (int b = 0; b < 5; ++b)
    if (b < nbytes) acc |= (uint64_t)bm[byte0 + b] << (8 * b);
  acc >>= sh;
  uint32_t r = (uint32_t)acc;
  if (nbits < 32) r &= (1u << nbits) - 1u;
  return r;
}

}  // namespace dfgpu
