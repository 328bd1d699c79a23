// microbench.cu — measures the B200 primitives the join / group-by designs depend on:
// random L2 atomics (RED.64 / RED.32), same-sector pairs, smem atomics, random 16-byte HBM reads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench scripts/microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__host__ __device__ inline uint64_t mix64(uint64_t x) { x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31; return x; }

template <int MODE>  // 0: one RED.64; 1: two RED.64 in separate arrays; 2: two RED.64 adjacent (same 16B); 3: one RED.32; 4: two RED.32 adjacent; 5: RED.64 + RED.32 adjacent
__global__ void red_kernel(unsigned long long* a, unsigned long long* b, uint64_t slots, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s = __umul64hi(mix64(i + 12345), slots);
    if (MODE == 0) atomicAdd(&a[s], (unsigned long long)i);
    if (MODE == 1) { atomicAdd(&a[s], (unsigned long long)i); atomicAdd(&b[s], 1ull); }
    if (MODE == 2) { atomicAdd(&a[2 * s], (unsigned long long)i); atomicAdd(&a[2 * s + 1], 1ull); }
    if (MODE == 3) atomicAdd((unsigned int*)a + s, (unsigned int)i);
    if (MODE == 4) { atomicAdd((unsigned int*)a + 2 * s, (unsigned int)i); atomicAdd((unsigned int*)a + 2 * s + 1, 1u); }
    if (MODE == 5) { atomicAdd(&a[2 * s], (unsigned long long)i); atomicAdd((unsigned int*)&a[2 * s + 1], 1u); }
  }
}
// streaming read of 16 B/row + the REDs (what the group-by kernel does)
__global__ void stream_red_kernel(const int4* in, unsigned long long* a, unsigned long long* b, uint64_t slots, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 v = in[i];
    uint64_t k = ((uint64_t)(uint32_t)v.y << 32) | (uint32_t)v.x;
    uint64_t s = __umul64hi(mix64(k), slots);
    atomicAdd(&a[s], ((unsigned long long)(uint32_t)v.w << 32) | (uint32_t)v.z);
    atomicAdd(&b[s], 1ull);
  }
}
// the group-by inner loop: stream {key,value}, find the key's slot (tag compare), 2 REDs; R rows per thread
__global__ void fill_rand_kernel(int4* in, int64_t n, uint64_t groups) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = mix64(i * 2 + 1) % groups, v = mix64(i * 2 + 2);
    in[i] = make_int4((int)k, (int)(k >> 32), (int)v, (int)(v >> 32));
  }
}
__global__ void fill_tags_kernel(unsigned long long* tags, uint64_t slots, uint64_t groups) {
  for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t s = __umul64hi(mix64(g), slots);
    while (atomicCAS(&tags[s], ~0ull, (unsigned long long)g) != ~0ull) { if (tags[s] == g) break; if (++s == slots) s = 0; }
  }
}
template <int R, int TAG>
__global__ void groupby_kernel(const int4* __restrict__ in, const unsigned long long* __restrict__ tags, unsigned long long* a, unsigned long long* b,
                               uint64_t slots, int64_t n) {
  for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * R; base < n; base += (int64_t)gridDim.x * blockDim.x * R) {
    int4 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = base + r < n ? in[base + r] : make_int4(0, 0, 0, 0);
    uint64_t s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { uint64_t k = ((uint64_t)(uint32_t)v[r].y << 32) | (uint32_t)v[r].x; s[r] = __umul64hi(mix64(k), slots); }
    if (TAG) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        uint64_t k = ((uint64_t)(uint32_t)v[r].y << 32) | (uint32_t)v[r].x;
        while (__ldcg(&tags[s[r]]) != k) { if (++s[r] == slots) s[r] = 0; }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) if (base + r < n) {
      atomicAdd(&a[s[r]], ((unsigned long long)(uint32_t)v[r].w << 32) | (uint32_t)v[r].z);
      atomicAdd(&b[s[r]], 1ull);
    }
  }
}
template <int W64>
__global__ void smem_atomic_kernel(int64_t n, int slots, unsigned long long* out) {
  extern __shared__ unsigned long long sm[];
  for (int i = threadIdx.x; i < slots; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t s = (uint32_t)(mix64(i) % (uint64_t)slots);
    if (W64) atomicAdd(&sm[s], (unsigned long long)i); else atomicAdd((unsigned int*)sm + s, (unsigned int)i);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}
__global__ void random_read16_kernel(const uint4* table, uint64_t slots, int64_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s = __umul64hi(mix64(i + 777), slots);
    uint4 v = __ldcg(&table[s]);
    acc += v.x + v.z;
  }
  if (acc == 0x1234567) out[0] = acc;
}
__global__ void random_read8_kernel(const uint2* table, uint64_t slots, int64_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t s = __umul64hi(mix64(i + 777), slots);
    uint2 v = __ldcg(&table[s]);
    acc += v.x;
  }
  if (acc == 0x1234567) out[0] = acc;
}
template <class F> float time_it(F f, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  const int64_t n = 1ll << 28;  // 268M ops
  unsigned long long *a, *b, *out; int4* in;
  CK(cudaMalloc(&a, 1ull << 30)); CK(cudaMalloc(&b, 1ull << 30)); CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&in, (size_t)n * 16));
  CK(cudaMemset(a, 0, 1ull << 30)); CK(cudaMemset(b, 0, 1ull << 30)); CK(cudaMemset(in, 1, (size_t)n * 16));
  const int grid = 148 * 8, blk = 256;
  for (uint64_t slots : {1ull << 22}) {
    printf("slots=%llu (%.0f MB of u64)\n", (unsigned long long)slots, slots * 8 / 1e6);
    float t;
    t = time_it([&] { red_kernel<0><<<grid, blk>>>(a, b, slots, n); }); printf("  1x RED.64            : %7.3f ms  %6.1f Gops/s\n", t, n / t / 1e6);
    t = time_it([&] { red_kernel<1><<<grid, blk>>>(a, b, slots, n); }); printf("  2x RED.64 sep arrays : %7.3f ms  %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { red_kernel<2><<<grid, blk>>>(a, b, slots, n); }); printf("  2x RED.64 adjacent   : %7.3f ms  %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { red_kernel<3><<<grid, blk>>>(a, b, slots, n); }); printf("  1x RED.32            : %7.3f ms  %6.1f Gops/s\n", t, n / t / 1e6);
    t = time_it([&] { red_kernel<4><<<grid, blk>>>(a, b, slots, n); }); printf("  2x RED.32 adjacent   : %7.3f ms  %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { red_kernel<5><<<grid, blk>>>(a, b, slots, n); }); printf("  RED.64+RED.32 adj    : %7.3f ms  %6.1f Grows/s\n", t, n / t / 1e6);
  }
  {
    const uint64_t groups = 1000000, slots = 4000000;
    unsigned long long* tags; CK(cudaMalloc(&tags, slots * 8)); CK(cudaMemset(tags, 0xFF, slots * 8));
    fill_rand_kernel<<<grid, blk>>>(in, n, groups); fill_tags_kernel<<<grid, blk>>>(tags, slots, groups); CK(cudaDeviceSynchronize());
    float t;
    t = time_it([&] { groupby_kernel<1, 0><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("groupby R=1 no-tag : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { groupby_kernel<1, 1><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("groupby R=1 tag    : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { groupby_kernel<4, 1><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("groupby R=4 tag    : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { groupby_kernel<4, 1><<<148 * 16, 128>>>(in, tags, a, b, slots, n); }); printf("groupby R=4 tag 128thr: %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { groupby_kernel<8, 1><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("groupby R=8 tag    : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
  }
  for (int slots : {2048, 8192, 16384}) {
    CK(cudaFuncSetAttribute(smem_atomic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(smem_atomic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float t = time_it([&] { smem_atomic_kernel<1><<<148, 1024, slots * 8>>>(n, slots, out); }); printf("smem atomic u64 slots=%d: %7.3f ms %6.1f Gops/s\n", slots, t, n / t / 1e6);
    t = time_it([&] { smem_atomic_kernel<0><<<148, 1024, slots * 8>>>(n, slots, out); }); printf("smem atomic u32 slots=%d: %7.3f ms %6.1f Gops/s\n", slots, t, n / t / 1e6);
  }
  for (uint64_t slots : {1ull << 22, 1ull << 24, 1ull << 25, 1ull << 26}) {
    float t = time_it([&] { random_read16_kernel<<<grid, blk>>>((const uint4*)a, slots, n, out); });
    printf("random 16B reads over %5.0f MB: %7.3f ms  %6.1f Gops/s  (%.0f GB/s of 32B sectors)\n", slots * 16 / 1e6, t, n / t / 1e6, n * 32.0 / t / 1e6);
    if (slots * 8 <= (1ull << 30)) { t = time_it([&] { random_read8_kernel<<<grid, blk>>>((const uint2*)a, slots, n, out); });
    printf("random  8B reads over %5.0f MB: %7.3f ms  %6.1f Gops/s\n", slots * 8 / 1e6, t, n / t / 1e6); }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
