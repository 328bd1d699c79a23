// batch.cu — transfers, bitmap utilities, concat and take (arrow-select `take` / `concat_batches`
// as called from joins/utils.rs:1371,1379 and hash_join/exec.rs:2705).
#include "batch.cuh"
#include "scan.cuh"

namespace dfgpu {

// ------------------------------------------------------------------------------------------
// bitmap kernels
// ------------------------------------------------------------------------------------------
__global__ void bitmap_or_copy_kernel(uint32_t* __restrict__ dst, int64_t dst_off, const uint8_t* __restrict__ src,
                                      int64_t src_off, int64_t n) {
  // one thread per destination 32-bit word
  int64_t first_word = dst_off >> 5, last_word = (dst_off + n - 1) >> 5;
  for (int64_t w = first_word + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w <= last_word;
       w += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = max(w << 5, dst_off), hi = min((w + 1) << 5, dst_off + n);  // dst bit range
    int nb = (int)(hi - lo);
    uint32_t bits = load_bits32(src, src_off + (lo - dst_off), nb);
    bits <<= (int)(lo - (w << 5));
    if (bits) atomicOr(&dst[w], bits);
  }
}

__global__ void bitmap_set_range_kernel(uint32_t* __restrict__ dst, int64_t dst_off, int64_t n) {
  int64_t first_word = dst_off >> 5, last_word = (dst_off + n - 1) >> 5;
  for (int64_t w = first_word + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w <= last_word;
       w += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = max(w << 5, dst_off), hi = min((w + 1) << 5, dst_off + n);
    int nb = (int)(hi - lo);
    uint32_t bits = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
    bits <<= (int)(lo - (w << 5));
    atomicOr(&dst[w], bits);
  }
}

__global__ void bitmap_popcount_kernel(const uint8_t* __restrict__ bm, int64_t off, int64_t n, unsigned long long* out) {
  unsigned long long local = 0;
  int64_t nchunks = (n + 31) / 32;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nchunks; c += (int64_t)gridDim.x * blockDim.x) {
    int nb = (int)min((int64_t)32, n - c * 32);
    local += __popc(load_bits32(bm, off + c * 32, nb));
  }
  unsigned long long tot = block_reduce_sum<256, unsigned long long>(local);
  if (threadIdx.x == 0 && tot) atomicAdd(out, tot);
}

void bitmap_or_copy(dfgpu_ctx* ctx, uint8_t* dst, int64_t dst_off, const uint8_t* src, int64_t src_off, int64_t n) {
  if (n <= 0) return;
  int64_t words = ((dst_off + n - 1) >> 5) - (dst_off >> 5) + 1;
  bitmap_or_copy_kernel<<<grid_for(words, 256), 256, 0, ctx->stream>>>((uint32_t*)dst, dst_off, src, src_off, n);
  DF_LAUNCH_CHECK(ctx);
}
void bitmap_set_range(dfgpu_ctx* ctx, uint8_t* dst, int64_t dst_off, int64_t n) {
  if (n <= 0) return;
  int64_t words = ((dst_off + n - 1) >> 5) - (dst_off >> 5) + 1;
  bitmap_set_range_kernel<<<grid_for(words, 256), 256, 0, ctx->stream>>>((uint32_t*)dst, dst_off, n);
  DF_LAUNCH_CHECK(ctx);
}
int64_t count_set_bits(dfgpu_ctx* ctx, const uint8_t* bm, int64_t bit_offset, int64_t n) {
  if (n <= 0 || !bm) return n > 0 ? n : 0;
  DevBuf cnt(ctx, 8);
  cnt.zero();
  bitmap_popcount_kernel<<<grid_for((n + 31) / 32, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(bm, bit_offset, n, cnt.as<unsigned long long>());
  DF_LAUNCH_CHECK(ctx);
  return (int64_t)read_scalar<unsigned long long>(ctx, cnt.as<unsigned long long>());
}

// ------------------------------------------------------------------------------------------
// transfers
// ------------------------------------------------------------------------------------------
DCol upload_column(dfgpu_ctx* ctx, const dfgpu_column& hc) {
  DCol d;
  d.type = hc.type; d.length = hc.length; d.null_count = hc.null_count;
  int w = type_width(hc.type);
  DF_CHECK(w >= 0, DFGPU_ERR_UNSUPPORTED, "unsupported column type");
  const bool has_valid = hc.validity != nullptr && hc.null_count != 0;
  if (hc.type == DFGPU_BOOL) {
    // copy the bytes covering [offset, offset+len); keep the residual bit offset
    int64_t b0 = hc.offset >> 3, b1 = (hc.offset + hc.length + 7) >> 3;
    size_t nbytes = (size_t)(b1 - b0);
    d.own_values = std::make_shared<DevBuf>(ctx, ((nbytes + 7) / 8) * 8 + 8);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_values->ptr, (const uint8_t*)hc.values + b0, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    d.values = d.own_values->ptr;
    d.offset = hc.offset & 7;
  } else {
    size_t nbytes = (size_t)hc.length * w;
    d.own_values = std::make_shared<DevBuf>(ctx, nbytes);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_values->ptr, (const char*)hc.values + hc.offset * w, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    d.values = d.own_values->ptr;
    d.offset = 0;
  }
  if (has_valid) {
    int64_t b0 = hc.offset >> 3, b1 = (hc.offset + hc.length + 7) >> 3;
    size_t nbytes = (size_t)(b1 - b0);
    d.own_validity = std::make_shared<DevBuf>(ctx, ((nbytes + 7) / 8) * 8 + 8);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_validity->ptr, hc.validity + b0, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    d.validity = d.own_validity->as<uint8_t>();
    d.offset = hc.offset & 7;  // BOOL values and validity share the same residual offset
  } else {
    d.validity = nullptr;
    if (hc.type != DFGPU_BOOL) d.offset = 0;
    d.null_count = 0;
  }
  return d;
}

DCol copy_column_device(dfgpu_ctx* ctx, const DCol& s) {
  DCol d;
  d.type = s.type; d.length = s.length; d.null_count = s.null_count;
  int w = type_width(s.type);
  int64_t res = s.offset & 7;
  if (s.type == DFGPU_BOOL) {
    int64_t b0 = s.offset >> 3, b1 = (s.offset + s.length + 7) >> 3;
    size_t nbytes = (size_t)(b1 - b0);
    d.own_values = std::make_shared<DevBuf>(ctx, ((nbytes + 7) / 8) * 8 + 8);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_values->ptr, (const uint8_t*)s.values + b0, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    d.values = d.own_values->ptr;
    d.offset = res;
  } else {
    size_t nbytes = (size_t)s.length * w;
    d.own_values = std::make_shared<DevBuf>(ctx, nbytes);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_values->ptr, s.values, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    d.values = d.own_values->ptr;
    d.offset = 0;
  }
  if (s.validity) {
    int64_t b0 = s.offset >> 3, b1 = (s.offset + s.length + 7) >> 3;
    size_t nbytes = (size_t)(b1 - b0);
    d.own_validity = std::make_shared<DevBuf>(ctx, ((nbytes + 7) / 8) * 8 + 8);
    if (nbytes) DF_CUDA(cudaMemcpyAsync(d.own_validity->ptr, s.validity + b0, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    d.validity = d.own_validity->as<uint8_t>();
    d.offset = res;
  }
  return d;
}

DCol slice_column(const DCol& c, int64_t start, int64_t len) {
  DCol d = c;
  d.length = len;
  d.null_count = (c.null_count == 0) ? 0 : -1;
  if (c.type == DFGPU_BOOL) {
    d.offset = c.offset + start;
  } else {
    d.values = (const char*)c.values + start * type_width(c.type);
    d.offset = c.validity ? c.offset + start : 0;
  }
  return d;
}

BatchPtr to_host_batch(dfgpu_ctx* ctx, const dfgpu_batch& dev) {
  BatchPtr hb(new dfgpu_batch());
  hb->ctx = ctx; hb->rows = dev.rows; hb->host = true;
  for (const DCol& c : dev.cols) {
    HCol h;
    h.type = c.type; h.length = c.length; h.null_count = c.null_count;
    // normalise to offset 0 on the host side: values copied from logical element 0
    if (c.type == DFGPU_BOOL) {
      // produce an offset-0 bitmap on device first if needed
      size_t nbytes = bitmap_bytes(c.length);
      h.values = std::make_shared<HostBuf>(nbytes ? ((nbytes + 7) / 8) * 8 : 8);
      if (c.length) {
        if ((c.offset & 7) == 0) {
          DF_CUDA(cudaMemcpyAsync(h.values->ptr, (const uint8_t*)c.values + (c.offset >> 3), nbytes, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
          DevBuf tmp(ctx, bitmap_alloc_bytes(c.length));
          tmp.zero();
          bitmap_or_copy(ctx, tmp.as<uint8_t>(), 0, (const uint8_t*)c.values, c.offset, c.length);
          DF_CUDA(cudaMemcpyAsync(h.values->ptr, tmp.ptr, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
      }
    } else {
      size_t nbytes = values_bytes(c.type, c.length);
      h.values = std::make_shared<HostBuf>(nbytes ? nbytes : 8);
      if (nbytes) DF_CUDA(cudaMemcpyAsync(h.values->ptr, c.values, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (c.validity && c.length) {
      size_t nbytes = bitmap_bytes(c.length);
      h.validity = std::make_shared<HostBuf>(((nbytes + 7) / 8) * 8);
      if ((c.offset & 7) == 0) {
        DF_CUDA(cudaMemcpyAsync(h.validity->ptr, c.validity + (c.offset >> 3), nbytes, cudaMemcpyDeviceToHost, ctx->stream));
      } else {
        DevBuf tmp(ctx, bitmap_alloc_bytes(c.length));
        tmp.zero();
        bitmap_or_copy(ctx, tmp.as<uint8_t>(), 0, c.validity, c.offset, c.length);
        DF_CUDA(cudaMemcpyAsync(h.validity->ptr, tmp.ptr, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
      }
      if (h.null_count < 0) h.null_count = c.length - count_set_bits(ctx, c.validity, c.offset, c.length);
    } else {
      h.null_count = 0;
    }
    hb->hcols.push_back(std::move(h));
  }
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  return hb;
}

// ------------------------------------------------------------------------------------------
// concat
// ------------------------------------------------------------------------------------------
DCol concat_columns(dfgpu_ctx* ctx, const std::vector<DCol>& parts, int type) {
  int64_t total = 0;
  bool any_valid = false;
  for (auto& p : parts) { total += p.length; any_valid |= (p.validity != nullptr); }
  if (parts.size() == 1) return parts[0];
  DCol d = alloc_col(ctx, type, total, any_valid);
  if (type == DFGPU_BOOL) d.own_values->zero();
  if (any_valid) d.own_validity->zero();
  int64_t pos = 0;
  int w = type_width(type);
  for (auto& p : parts) {
    if (p.length == 0) continue;
    if (type == DFGPU_BOOL) {
      bitmap_or_copy(ctx, (uint8_t*)d.own_values->ptr, pos, (const uint8_t*)p.values, p.offset, p.length);
    } else {
      DF_CUDA(cudaMemcpyAsync((char*)d.own_values->ptr + pos * w, p.values, (size_t)p.length * w, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (any_valid) {
      if (p.validity) bitmap_or_copy(ctx, (uint8_t*)d.own_validity->ptr, pos, p.validity, p.offset, p.length);
      else bitmap_set_range(ctx, (uint8_t*)d.own_validity->ptr, pos, p.length);
    }
    pos += p.length;
  }
  d.null_count = any_valid ? -1 : 0;
  return d;
}

// ------------------------------------------------------------------------------------------
// take
// ------------------------------------------------------------------------------------------
constexpr uint32_t kNullIdx = 0xFFFFFFFFu;

template <class T>
__global__ void take_kernel(const T* __restrict__ src, const uint8_t* __restrict__ src_valid, int64_t src_voff,
                            const uint32_t* __restrict__ idx, int64_t n, T* __restrict__ out, uint32_t* __restrict__ out_valid) {
  // grid-stride over whole warps so that each warp owns 32 consecutive outputs (one validity word)
  int64_t nwarp_items = (n + 31) / 32;
  int lane = threadIdx.x & 31;
  for (int64_t wi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; wi < nwarp_items;
       wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t i = wi * 32 + lane;
    bool ok = false;
    T v = T();
    if (i < n) {
      uint32_t ix = idx[i];
      if (ix != kNullIdx) {
        ok = src_valid ? bit_get(src_valid, src_voff + ix) : true;
        v = src[ix];
      }
      out[i] = v;
    }
    if (out_valid) {
      uint32_t word = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) out_valid[wi] = word;
    }
  }
}

__global__ void take_bool_kernel(const uint8_t* __restrict__ src, int64_t src_off, const uint8_t* __restrict__ src_valid, int64_t src_voff,
                                 const uint32_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ out, uint32_t* __restrict__ out_valid) {
  int64_t nwarp_items = (n + 31) / 32;
  int lane = threadIdx.x & 31;
  for (int64_t wi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; wi < nwarp_items;
       wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t i = wi * 32 + lane;
    bool ok = false, v = false;
    if (i < n) {
      uint32_t ix = idx[i];
      if (ix != kNullIdx) {
        ok = src_valid ? bit_get(src_valid, src_voff + ix) : true;
        v = bit_get(src, src_off + ix);
      }
    }
    uint32_t vw = __ballot_sync(0xffffffffu, v);
    uint32_t ow = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) { out[wi] = vw; if (out_valid) out_valid[wi] = ow; }
  }
}

struct alignas(16) B16 { uint64_t a, b; };

DCol take_column(dfgpu_ctx* ctx, const DCol& src, const uint32_t* idx, int64_t n, bool idx_may_be_null) {
  bool need_valid = idx_may_be_null || src.validity != nullptr;
  DCol d = alloc_col(ctx, src.type, n, need_valid);
  if (n == 0) return d;
  uint32_t* ov = need_valid ? d.own_validity->as<uint32_t>() : nullptr;
  int grid = grid_for(n, 256, kNumSMs * 16);
  if (src.type == DFGPU_BOOL) {
    take_bool_kernel<<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src.values, src.offset, src.validity, src.offset, idx, n,
                                                   d.own_values->as<uint32_t>(), ov);
  } else {
    switch (type_width(src.type)) {
      case 1: take_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)src.values, src.validity, src.offset, idx, n, d.own_values->as<uint8_t>(), ov); break;
      case 2: take_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t*)src.values, src.validity, src.offset, idx, n, d.own_values->as<uint16_t>(), ov); break;
      case 4: take_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>((const uint32_t*)src.values, src.validity, src.offset, idx, n, d.own_values->as<uint32_t>(), ov); break;
      case 8: take_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>((const uint64_t*)src.values, src.validity, src.offset, idx, n, d.own_values->as<uint64_t>(), ov); break;
      case 16: take_kernel<B16><<<grid, 256, 0, ctx->stream>>>((const B16*)src.values, src.validity, src.offset, idx, n, d.own_values->as<B16>(), ov); break;
      default: throw Error(DFGPU_ERR_UNSUPPORTED, "take: unsupported width");
    }
  }
  DF_LAUNCH_CHECK(ctx);
  d.null_count = need_valid ? -1 : 0;
  return d;
}

DCol null_column(dfgpu_ctx* ctx, int type, int64_t n) {
  DCol d = alloc_col(ctx, type, n, true);
  d.own_values->zero();
  d.own_validity->zero();
  d.null_count = n;
  return d;
}


// ------------------------------------------------------------------------------------------
// flag-bitmap compaction (count -> single-block scan -> emit)
// ------------------------------------------------------------------------------------------
constexpr int kCompactThreads = 256;
// stream compaction of row indices by a flag bitmap (want_set selects set or clear bits)
__global__ void __launch_bounds__(kCompactThreads) flags_count_kernel(const uint32_t* __restrict__ words, int64_t n, int want_set, uint64_t* __restrict__ tile_sums) {
  int64_t i = (int64_t)blockIdx.x * kCompactThreads + threadIdx.x;  // one word per thread
  int64_t nw = (n + 31) / 32;
  uint32_t c = 0;
  if (i < nw) {
    uint32_t w = words[i];
    if (!want_set) w = ~w;
    int rem = (int)min((int64_t)32, n - i * 32);
    if (rem < 32) w &= (1u << rem) - 1u;
    c = __popc(w);
  }
  uint64_t tot = block_reduce_sum<kCompactThreads, uint64_t>((uint64_t)c);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kCompactThreads) flags_emit_kernel(const uint32_t* __restrict__ words, int64_t n, int want_set, const uint64_t* __restrict__ tile_offsets, uint32_t* __restrict__ out_idx) {
  int64_t i = (int64_t)blockIdx.x * kCompactThreads + threadIdx.x;
  int64_t nw = (n + 31) / 32;
  uint32_t w = 0;
  if (i < nw) {
    w = words[i];
    if (!want_set) w = ~w;
    int rem = (int)min((int64_t)32, n - i * 32);
    if (rem < 32) w &= (1u << rem) - 1u;
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan<kCompactThreads, uint32_t>(__popc(w), &tot);
  uint64_t pos = tile_offsets[blockIdx.x] + ex;
  while (w) {
    int b = __ffs(w) - 1;
    w &= w - 1;
    out_idx[pos++] = (uint32_t)(i * 32 + b);
  }
}

__global__ void iota_kernel(uint32_t* out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)i;
}

int64_t compact_count(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, DevBuf* tiles) {
  if (n <= 0) return 0;
  int64_t nw = (n + 31) / 32;
  int64_t ntiles = (nw + kCompactThreads - 1) / kCompactThreads;
  tiles->alloc(ctx, (size_t)(ntiles + 1) * 8);
  flags_count_kernel<<<(int)ntiles, kCompactThreads, 0, ctx->stream>>>(words, n, want_set, tiles->as<uint64_t>());
  DF_LAUNCH_CHECK(ctx);
  scan_tiles_kernel<1024><<<1, 1024, 0, ctx->stream>>>(tiles->as<uint64_t>(), ntiles, tiles->as<uint64_t>() + ntiles);
  DF_LAUNCH_CHECK(ctx);
  return (int64_t)read_scalar<uint64_t>(ctx, tiles->as<uint64_t>() + ntiles);
}
void compact_emit(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, const DevBuf& tiles, uint32_t* dst) {
  if (n <= 0) return;
  int64_t nw = (n + 31) / 32;
  int64_t ntiles = (nw + kCompactThreads - 1) / kCompactThreads;
  flags_emit_kernel<<<(int)ntiles, kCompactThreads, 0, ctx->stream>>>(words, n, want_set, tiles.as<uint64_t>(), dst);
  DF_LAUNCH_CHECK(ctx);
}
int64_t compact_flag_indices(dfgpu_ctx* ctx, const uint32_t* words, int64_t n, int want_set, DevBuf* out_idx) {
  DevBuf tiles;
  int64_t total = compact_count(ctx, words, n, want_set, &tiles);
  if (total == 0) return 0;
  out_idx->alloc(ctx, (size_t)total * 4);
  compact_emit(ctx, words, n, want_set, tiles, out_idx->as<uint32_t>());
  return total;
}
void fill_iota(dfgpu_ctx* ctx, uint32_t* out, int64_t n) {
  if (n <= 0) return;
  iota_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(out, n);
  DF_LAUNCH_CHECK(ctx);
}

}  // namespace dfgpu
