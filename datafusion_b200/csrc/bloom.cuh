// bloom.cuh — the split-block membership filter shared by the fused pipeline (pipeline.cu) and the stand-alone hash join (hash_join.cu)
#pragma once
#include "common.cuh"

namespace dfgpu {

// Membership filter = split-block Bloom filter: one 64-bit block per key (two 32-bit words, two probe bits in each), 16 bits per key.
// Block and bit positions come from a 32-bit multiplicative hash of both key halves — the filter is probed for EVERY scanned row, so
// its cost is counted in instructions: ~12 integer ops here against ~45 for mix64 + a 64-bit fastrange + four 64-bit shifts.
struct BloomPos { uint32_t block, t; };   // t: 20 hash bits = four 5-bit probe positions (two per 32-bit word)
__device__ __forceinline__ BloomPos bloom_pos(uint64_t key, uint64_t blocks) {
  uint32_t h1 = ((uint32_t)key ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
  h1 ^= h1 >> 15;
  BloomPos p;
  p.block = __umulhi(h1, (uint32_t)blocks);
  p.t = (h1 * 0xC2B2AE35u) >> 12;                           // the well-mixed upper 20 bits of a second multiply
  return p;
}
__device__ __forceinline__ unsigned long long bloom_mask(uint32_t t) {
  const uint32_t m0 = (1u << (t & 31)) | (1u << ((t >> 5) & 31)), m1 = (1u << ((t >> 10) & 31)) | (1u << ((t >> 15) & 31));
  return ((unsigned long long)m1 << 32) | m0;
}
__device__ __forceinline__ void bloom_set(unsigned long long* bloom, uint64_t blocks, uint64_t key) {
  const BloomPos p = bloom_pos(key, blocks);
  atomicOr(&bloom[p.block], bloom_mask(p.t));
}
__device__ __forceinline__ bool bloom_test(unsigned long long w, uint32_t t) {
  const uint32_t m0 = (1u << (t & 31)) | (1u << ((t >> 5) & 31)), m1 = (1u << ((t >> 10) & 31)) | (1u << ((t >> 15) & 31));
  return ((uint32_t)w & m0) == m0 && ((uint32_t)(w >> 32) & m1) == m1;
}

}  // namespace dfgpu
