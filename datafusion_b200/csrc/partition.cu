// partition.cu — the local pass of the hash exchange: RepartitionExec / BatchPartitioner::Hash
// (reference physical-plan/src/repartition/mod.rs:618-648, 1097-1145; partition_indices :895;
// partition_grouped_take :1237).  On 8 GPUs this pass feeds one NCCL all-to-all; rows keep their
// input order inside every partition (the reference's per-partition `take` is order preserving).
//
// round-1 implementation: partition id per row -> one flag bitmap per partition (warp ballots) ->
// per-partition ordered index compaction into one permutation -> one gather per column.
#include "batch.cuh"
#include "scan.cuh"

namespace dfgpu {

constexpr int kMaxPartKeys = 4;
struct PartKeys {
  int n;
  const void* ptr[kMaxPartKeys];
  const uint8_t* valid[kMaxPartKeys];
  int64_t voff[kMaxPartKeys];
  int width[kMaxPartKeys];
};

__device__ __forceinline__ uint64_t exchange_hash(const PartKeys& k, int64_t row) {
  // create_hashes with the repartition seed (repartition/mod.rs:650, 1126-1130): first column hashed
  // with the seed, later columns re-seeded with the running hash; NULLs leave the running hash untouched
  uint64_t h = 0;
  bool first = true;
#pragma unroll
  for (int c = 0; c < kMaxPartKeys; ++c) {
    if (c >= k.n) break;
    if (k.valid[c] && !bit_get(k.valid[c], k.voff[c] + row)) continue;
    uint64_t v;
    switch (k.width[c]) {
      case 1: v = ((const uint8_t*)k.ptr[c])[row]; break;
      case 2: v = ((const uint16_t*)k.ptr[c])[row]; break;
      case 4: v = ((const uint32_t*)k.ptr[c])[row]; break;
      default: v = ((const uint64_t*)k.ptr[c])[row]; break;
    }
    h = first ? hash_u64(v, kSeedExchange) : hash_combine(h, v);
    first = false;
  }
  return h;
}

__global__ void __launch_bounds__(256) partition_flags_kernel(PartKeys k, int64_t n, int n_parts, uint32_t* __restrict__ flag_words /* [n_parts][nw] */) {
  const int64_t nw = (n + 31) / 32;
  const int lane = threadIdx.x & 31;
  for (int64_t wi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; wi < nw; wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    int64_t row = wi * 32 + lane;
    int pid = -1;
    if (row < n) pid = (int)(exchange_hash(k, row) % (uint64_t)n_parts);  // hash % n (repartition/mod.rs:875-935)
    for (int p = 0; p < n_parts; ++p) {
      uint32_t w = __ballot_sync(0xffffffffu, pid == p);
      if (lane == 0) flag_words[(int64_t)p * nw + wi] = w;
    }
  }
}

}  // namespace dfgpu

using namespace dfgpu;

extern "C" int dfgpu_hash_partition_device(dfgpu_ctx* ctx, const dfgpu_column* cols, int32_t n_cols, const int32_t* key_cols, int32_t n_keys,
                                           int32_t n_parts, dfgpu_batch** out, int64_t* part_offsets_host) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && cols && key_cols && out && part_offsets_host, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(n_keys >= 1 && n_keys <= kMaxPartKeys, DFGPU_ERR_UNSUPPORTED, "hash partition: 1..4 key columns");
  DF_CHECK(n_parts >= 1 && n_parts <= 1024, DFGPU_ERR_INVALID, "hash partition: 1..1024 partitions");
  set_device(ctx);
  std::vector<DCol> v;
  for (int i = 0; i < n_cols; ++i) v.push_back(device_view(cols[i]));
  const int64_t n = n_cols ? v[0].length : 0;
  DF_CHECK(n < 0xFFFFFFFFll, DFGPU_ERR_UNSUPPORTED, "hash partition: < 2^32-1 rows per call");
  PartKeys pk;
  memset(&pk, 0, sizeof(pk));
  pk.n = n_keys;
  for (int c = 0; c < n_keys; ++c) {
    DF_CHECK(key_cols[c] >= 0 && key_cols[c] < n_cols, DFGPU_ERR_INVALID, "key column out of range");
    const DCol& col = v[key_cols[c]];
    int w = type_width(col.type);
    DF_CHECK(w >= 1 && w <= 8, DFGPU_ERR_UNSUPPORTED, "hash partition: key must be a fixed-width type of <= 64 bits");
    pk.ptr[c] = col.values; pk.valid[c] = col.validity; pk.voff[c] = col.offset; pk.width[c] = w;
  }
  BatchPtr b(new dfgpu_batch());
  b->ctx = ctx; b->rows = n; b->host = false;
  for (int p = 0; p <= n_parts; ++p) part_offsets_host[p] = 0;
  if (n > 0) {
    const int64_t nw = (n + 31) / 32;
    DevBuf flags(ctx, (size_t)n_parts * nw * 4), perm(ctx, (size_t)n * 4);
    partition_flags_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(pk, n, n_parts, flags.as<uint32_t>());
    DF_LAUNCH_CHECK(ctx);
    int64_t pos = 0;
    for (int p = 0; p < n_parts; ++p) {
      DevBuf tiles;
      const uint32_t* w = flags.as<uint32_t>() + (int64_t)p * nw;
      int64_t cnt = compact_count(ctx, w, n, 1, &tiles);
      part_offsets_host[p] = pos;
      if (cnt) compact_emit(ctx, w, n, 1, tiles, perm.as<uint32_t>() + pos);
      pos += cnt;
    }
    part_offsets_host[n_parts] = pos;
    DF_CHECK(pos == n, DFGPU_ERR_CUDA, "hash partition: internal row count mismatch");
    for (int i = 0; i < n_cols; ++i) b->cols.push_back(take_column(ctx, v[i], perm.as<uint32_t>(), n, false));
  } else {
    for (int i = 0; i < n_cols; ++i) b->cols.push_back(alloc_col(ctx, v[i].type, 0, false));
  }
  *out = b.release();
  DF_API_END
}
