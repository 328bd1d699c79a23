// scan.cuh — block-level scan / reduce primitives (warp shuffles + one smem round trip).
#pragma once
#include "common.cuh"

namespace dfgpu {

// exclusive scan of one value per thread across a block of NT threads (NT multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block total (valid in all threads).
template <int NT, class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total) {
  __shared__ T warp_sums[NT / 32 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  __syncthreads();  // protect warp_sums reuse across calls
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < NT / 32 ? warp_sums[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T n = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += n;
    }
    if (lane < NT / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
    if (lane == NT / 32 - 1) warp_sums[NT / 32] = wi;
  }
  __syncthreads();
  *total = warp_sums[NT / 32];
  return warp_sums[warp] + incl - v;
}

template <int NT, class T>
__device__ __forceinline__ T block_reduce_sum(T v) {
  __shared__ T red[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  T r = T(0);
  if (warp == 0) {
    r = lane < NT / 32 ? red[lane] : T(0);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
  }
  return r;  // valid in warp 0
}

// Exclusive scan of `n` uint64 tile sums, in place, one block per row: block r scans sums[r*n .. (r+1)*n) and
// writes that row's total to total_out[r] (a <<<1, NT>>> launch is the plain device-wide scan of n values).
template <int NT>
__global__ void scan_tiles_kernel(uint64_t* __restrict__ sums, int64_t n, uint64_t* __restrict__ total_out) {
  constexpr int IPT = 8;   // consecutive items per thread and round: 8192 items per block-wide scan
  sums += (int64_t)blockIdx.x * n;
  total_out += blockIdx.x;
  uint64_t carry = 0;
  for (int64_t base = 0; base < n; base += (int64_t)NT * IPT) {
    const int64_t i0 = base + (int64_t)threadIdx.x * IPT;
    uint64_t v[IPT], sum = 0;
#pragma unroll
    for (int k = 0; k < IPT; ++k) { v[k] = i0 + k < n ? sums[i0 + k] : 0; sum += v[k]; }
    uint64_t tot;
    uint64_t ex = carry + block_exclusive_scan<NT, uint64_t>(sum, &tot);
#pragma unroll
    for (int k = 0; k < IPT; ++k) { if (i0 + k < n) sums[i0 + k] = ex; ex += v[k]; }
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// ---- single-pass ordered compaction support: tile descriptors + decoupled look-back ----
#define kDescAgg (1ull << 62)
#define kDescPrefix (2ull << 62)
#define kDescMask ((1ull << 62) - 1ull)

// decoupled look-back over tile descriptors (status in the top 2 bits, value below); called by warp 0.
// Returns the exclusive prefix of `tot` for `tile` and publishes this tile's inclusive prefix.
__device__ __forceinline__ unsigned long long tile_lookback(int64_t tile, uint32_t tot, unsigned long long* tile_desc) {
  const int lane = threadIdx.x & 31;
  unsigned long long exclusive = 0;
  if (tile == 0) {
    if (lane == 0) atomicExch(&tile_desc[0], kDescPrefix | (unsigned long long)tot);
    return 0;
  }
  if (lane == 0) atomicExch(&tile_desc[tile], kDescAgg | (unsigned long long)tot);
  int64_t look = tile - 1;
  while (true) {
    const int64_t idx = look - lane;
    unsigned long long d = idx >= 0 ? *(volatile unsigned long long*)&tile_desc[idx] : kDescPrefix;  // virtual prefix 0 before tile 0
    const unsigned st = (unsigned)(d >> 62);
    const unsigned invalid = __ballot_sync(0xffffffffu, st == 0);
    const unsigned prefix = __ballot_sync(0xffffffffu, st == 2);
    const int first_prefix = prefix ? __ffs(prefix) - 1 : 32;
    const int first_invalid = invalid ? __ffs(invalid) - 1 : 32;
    if (first_invalid < first_prefix) continue;  // a predecessor in the window has not published yet: spin
    unsigned long long contrib = (lane <= first_prefix) ? (d & kDescMask) : 0ull;
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, dd);
    exclusive += contrib;
    if (first_prefix < 32) break;
    look -= 32;
  }
  if (lane == 0) atomicExch(&tile_desc[tile], kDescPrefix | (exclusive + (unsigned long long)tot));
  return exclusive;
}


}  // namespace dfgpu
