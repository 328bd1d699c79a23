/* gen.h — counter-based synthetic data generators, bit-identical to the device versions in
 * datafusion_b200/csrc/common.cuh (splitmix64_at, mix64) and context.cu (perm_bijection).
 * TEST INFRASTRUCTURE ONLY (oracle/): never linked into libdfgpu.so. */
#ifndef ORACLE_GEN_H
#define ORACLE_GEN_H
#include <stdint.h>

static inline uint64_t o_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
static inline uint64_t o_splitmix64_at(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t o_perm_bijection(uint64_t i, uint64_t n, uint64_t seed) {
  int bits = 1;
  while ((1ull << bits) < n) ++bits;
  if (bits & 1) ++bits;
  const int half = bits / 2;
  const uint64_t mask = (1ull << half) - 1ull;
  uint64_t x = i;
  do {
    uint64_t l = x >> half, r = x & mask;
    for (int round = 0; round < 4; ++round) {
      uint64_t f = o_mix64(r + seed * 0x9E3779B97F4A7C15ull + (uint64_t)round * 0xD1B54A32D192ED03ull) & mask;
      uint64_t nl = r, nr = l ^ f;
      l = nl; r = nr;
    }
    x = (l << half) | r;
  } while (x >= n);
  return x;
}
enum { O_GEN_SEQ = 0, O_GEN_UNIFORM = 1, O_GEN_SPLITMIX = 2, O_GEN_PERM = 3, O_GEN_SPARSE_OF = 4 };
static inline int64_t o_gen_value(int kind, uint64_t seed, int64_t a, int64_t b, uint64_t idx) {
  switch (kind) {
    case O_GEN_SEQ: return a + (int64_t)idx;
    case O_GEN_UNIFORM: return a + (int64_t)(o_splitmix64_at(seed, idx) % (uint64_t)b);
    case O_GEN_SPLITMIX: return (int64_t)o_splitmix64_at(seed, idx);
    case O_GEN_PERM: return a + (int64_t)o_perm_bijection(idx, (uint64_t)b, seed);
    case O_GEN_SPARSE_OF: return (int64_t)o_splitmix64_at(seed, o_splitmix64_at((uint64_t)a, idx) % (uint64_t)b);
    default: return 0;
  }
}
#endif
