// microbench_gran.cu — does cudaLimitMaxL2FetchGranularity change the DRAM cost of a random 16 B table lookup?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mbg scripts/microbench_gran.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31; return x; }
template <int R>
__global__ void rand16(const uint4* __restrict__ t, uint64_t slots, int64_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * R; i < n; i += (int64_t)gridDim.x * blockDim.x * R) {
    uint4 v[R];
#pragma unroll
    for (int k = 0; k < R; ++k) v[k] = __ldcg(t + __umul64hi(mix64(i + k + 12345), slots));
#pragma unroll
    for (int k = 0; k < R; ++k) acc += v[k].x + v[k].z;
  }
  if (acc == 0x1234567) out[0] = acc;
}
int main() {
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("default fetch granularity %zu\n", g);
  const int64_t n = 100000000;
  unsigned long long* out; cudaMalloc(&out, 8);
  for (size_t mb : {100, 400, 1600}) {
    const uint64_t slots = mb * 1000000ull / 16;
    uint4* t; cudaMalloc(&t, slots * 16); cudaMemset(t, 1, slots * 16);
    for (size_t gran : {128, 64, 32}) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
      cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      float best = 1e9;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(a); rand16<4><<<148 * 16, 256>>>(t, slots, n, out); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it && ms < best) best = ms;
      }
      printf("table %zu MB gran set %zu (rc %d, now %zu): %.3f ms  %.1f Gop/s\n", mb, gran, (int)e, g, best, n / best * 1e-6);
    }
    cudaFree(t);
  }
  return 0;
}
