// microbench_gb.cu — what bounds the hash group-by inner loop on B200?  (1B-row C3 shape scaled to 268M rows)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__host__ __device__ inline uint64_t mix64(uint64_t x) { x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31; return x; }
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r; unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__global__ void fill_rand_kernel(int4* in, int64_t n, uint64_t groups) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t k = mix64(i * 2 + 1) % groups, v = mix64(i * 2 + 2);
    in[i] = make_int4((int)k, (int)(k >> 32), (int)v, (int)(v >> 32));
  }
}
__global__ void fill_tags_kernel(unsigned long long* tags, int stride, uint64_t slots, uint64_t groups) {
  for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t s = __umul64hi(mix64(g), slots);
    while (atomicCAS(&tags[s * stride], ~0ull, (unsigned long long)g) != ~0ull) { if (tags[s * stride] == g) break; if (++s == slots) s = 0; }
  }
}
// MODE 0: SoA tags + 2 REDs (the current kernel)   1: tag lookup only   2: AoS 32B slot {tag,sum,cnt,pad}: tag load + 2 REDs in one sector
// MODE 3: SoA, no tag (direct)   4: AoS, one RED only (sum) + count folded as second RED same sector skipped (lower bound)
template <int MODE, int EF>
__global__ void gb_kernel(const int4* __restrict__ in, unsigned long long* tags, unsigned long long* a, unsigned long long* b, uint64_t slots, int64_t n) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int4 v = EF ? ld_stream(in + i) : in[i];
    uint64_t k = ((uint64_t)(uint32_t)v.y << 32) | (uint32_t)v.x;
    unsigned long long val = ((unsigned long long)(uint32_t)v.w << 32) | (uint32_t)v.z;
    uint64_t s = __umul64hi(mix64(k), slots);
    if (MODE == 0) { while (__ldcg(&tags[s]) != k) { if (++s == slots) s = 0; } atomicAdd(&a[s], val); atomicAdd(&b[s], 1ull); }
    if (MODE == 1) { while (__ldcg(&tags[s]) != k) { if (++s == slots) s = 0; } acc += s; }
    if (MODE == 2) { while (__ldcg(&tags[s * 4]) != k) { if (++s == slots) s = 0; } atomicAdd(&tags[s * 4 + 1], val); atomicAdd(&tags[s * 4 + 2], 1ull); }
    if (MODE == 3) { atomicAdd(&a[s], val); atomicAdd(&b[s], 1ull); }
    if (MODE == 4) { while (__ldcg(&tags[s * 4]) != k) { if (++s == slots) s = 0; } atomicAdd(&tags[s * 4 + 1], val); }
  }
  if (acc == 0x12345) a[0] = acc;
}
template <class F> float time_it(F f, int reps = 3) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  const int64_t n = 1ll << 28;
  const uint64_t groups = 1000000;
  unsigned long long *a, *b, *tags, *aos; int4* in;
  CK(cudaMalloc(&a, 1ull << 28)); CK(cudaMalloc(&b, 1ull << 28)); CK(cudaMalloc(&tags, 1ull << 28)); CK(cudaMalloc(&aos, 1ull << 29)); CK(cudaMalloc(&in, (size_t)n * 16));
  const int grid = 148 * 8, blk = 256;
  fill_rand_kernel<<<grid, blk>>>(in, n, groups);
  for (uint64_t slots : {2000000ull, 4000000ull, 8000000ull}) {
    CK(cudaMemset(tags, 0xFF, slots * 8)); CK(cudaMemset(aos, 0xFF, slots * 32)); CK(cudaMemset(a, 0, slots * 8)); CK(cudaMemset(b, 0, slots * 8));
    fill_tags_kernel<<<grid, blk>>>(tags, 1, slots, groups); fill_tags_kernel<<<grid, blk>>>(aos, 4, slots, groups); CK(cudaDeviceSynchronize());
    printf("slots=%llu: SoA tags %.0f MB + acc 2x%.0f MB | AoS %.0f MB\n", (unsigned long long)slots, slots * 8 / 1e6, slots * 8 / 1e6, slots * 32 / 1e6);
    float t;
    t = time_it([&] { gb_kernel<3, 0><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("  no tag, 2 RED            : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<1, 0><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("  tag lookup only          : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<0, 0><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("  SoA tag + 2 RED          : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<0, 1><<<grid, blk>>>(in, tags, a, b, slots, n); }); printf("  SoA tag + 2 RED evict1st : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<2, 0><<<grid, blk>>>(in, aos, a, b, slots, n); }); printf("  AoS32 tag + 2 RED        : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<2, 1><<<grid, blk>>>(in, aos, a, b, slots, n); }); printf("  AoS32 tag + 2 RED evict1 : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
    t = time_it([&] { gb_kernel<4, 1><<<grid, blk>>>(in, aos, a, b, slots, n); }); printf("  AoS32 tag + 1 RED evict1 : %7.3f ms %6.1f Grows/s\n", t, n / t / 1e6);
  }
  return 0;
}
