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

// Device-wide exclusive scan of `n` uint64 tile sums, in place, by a single block (n is the
// number of tiles: <= a few hundred thousand).  total written to *total_out.
template <int NT>
__global__ void scan_tiles_kernel(uint64_t* __restrict__ sums, int64_t n, uint64_t* __restrict__ total_out) {
  uint64_t carry = 0;
  for (int64_t base = 0; base < n; base += NT) {
    int64_t i = base + threadIdx.x;
    uint64_t v = i < n ? sums[i] : 0;
    uint64_t tot;
    uint64_t ex = block_exclusive_scan<NT, uint64_t>(v, &tot);
    if (i < n) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

}  // namespace dfgpu
