// pipeline.cu — fused pipelines: FilterExec -> HashJoinExec probe side(s) -> {join build | AggregateExec | output}.
//
// Reference path being replaced (§8f rank 3):
//   FilterExecStream::poll_next                 physical-plan/src/filter.rs:1364-1445
//   HashJoinStream::process_probe_batch         physical-plan/src/joins/hash_join/stream.rs:740-1000
//   lookup_join_hashmap / equal_rows_arr        stream.rs:396-438, joins/utils.rs:2191-2257
//   adjust_indices_by_join_type (RightSemi/Anti) joins/utils.rs:1432-1490
//   AggregateHashTable::aggregate_batch_inner   aggregates/aggregate_hash_table/common.rs:205-236
//   PrimitiveGroupsAccumulator / CountGroupsAccumulator  prim_op.rs:41-195, count.rs:631-780
//   dynamic filter pushdown (bounds + membership) joins/hash_join/shared_bounds.rs, partitioned_hash_eval.rs
// and the plan shape it serves: sqllogictest/test_files/tpch/plans/q3.slt.part:60-76.
//
// B200 design.  The reference keeps a pipeline's intermediates in CPU cache by streaming 8192-row batches through
// the operators; on the GPU the same effect needs ONE kernel per pipeline: a thread block walks 1024-row tiles of the
// probe-side table, evaluates the predicate from coalesced column loads, probes the build-side structures and feeds
// the sink, so every input byte crosses HBM once and no filtered copy / join output / projection is materialised.
//   * a build side is a `dfgpu_lookup`: an open-addressing table of fixed-stride records
//       {key:u64 | payload:u64 | accumulator words...}   (exact key in the record: no equal_rows re-check),
//     or — for key sets over a dense range — a bitmap (the reference's ArrayMap idea at one bit per key).
//   * a blocked Bloom filter (one 64-bit block per lookup, 16 bits per key, 4 probes) sits in front of tables that
//     exceed L2: it is the device form of the reference's dynamic filter pushdown (membership test pushed into the
//     probe-side scan) and turns nine out of ten random DRAM accesses of a low-hit-rate join into L2 hits.
//   * when the GROUP BY keys are the join key plus build-side columns (TPC-H Q3: l_orderkey, o_orderdate,
//     o_shippriority) the group id IS the build row, so the accumulators live inside the matched record and an
//     update is one or two RED operations on the sector the probe just fetched — no second hash table.
#include "batch.cuh"
#include "scan.cuh"
#include "expr.cuh"
#include "expr_dev.cuh"
#include "expr_dec.cuh"
#include "bloom.cuh"
#include <climits>
#include <array>

namespace dfgpu {

std::vector<dfgpu_column> arrow_to_columns(const ArrowArray* batch, const ArrowSchema* schema);  // arrow_io.cu

constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kMaxPipeCols = 16, kMaxStages = 3, kMaxPipeAggs = 4, kMaxExt = 8, kMaxTerms = 4, kPoolNodes = 56, kMaxBuildPay = 8;
constexpr int kPipeThreads = 256, kPipeItems = 4, kPipeTile = kPipeThreads * kPipeItems;
enum LookupMode : int { LK_HASH = 0, LK_BITMAP = 1 };
enum SinkKind : int { SINK_NONE = 0, SINK_COUNT = 1, SINK_BUILD = 2, SINK_AGG = 3, SINK_OUTPUT = 4, SINK_OUTPUT_ANY = 5 /* row order unspecified */,
                      SINK_PACK = 6 /* build sink, table size unknown: {key, payload} records to a staging buffer, inserted afterwards */ };
constexpr int kStageMaybe = 3;   // DFGPU_STAGE_MAYBE
constexpr int kPipeVarDefault = 43;   // pipe_kernel's VAR when DFGPU_PIPE_VAR is not set

struct LookupDev {
  int mode, stride /* 8-byte words per record */, has_payload, pad;
  unsigned long long* recs; uint64_t cap;
  unsigned long long* bloom; uint64_t bloom_blocks;
  uint32_t* coarse; uint64_t coarse_words;   // optional first level (4 bits per key, L2-resident) in front of a filter that exceeds L2
  uint32_t* bits; uint64_t kmin, ksize;
};
struct ColRef { const void* ptr; const uint8_t* valid; int64_t voff; int width, sgn, vec /* base pointer 16-byte aligned: 128-bit loads allowed */, pad; };
struct StageDev { int kind, key_col; LookupDev lk; };
struct ExtDef { int stage, shift, width, type; };
struct AggDef { int func, cls, word, nn_word, start, n, small, pad; };
struct PipeParams {
  int n_cols, hints; ColRef col[kMaxPipeCols];
  int pred_mode /* 0 none, 1 conjunction of col <cmp> literal, 2 interpreter */, n_terms, pred_start, pred_n, pred_small, first_hash /* first stage backed by a hash table, -1 = none */;
  int term_col[kMaxTerms], term_op[kMaxTerms], term_uns[kMaxTerms]; long long term_lit[kMaxTerms];
  int n_stages; StageDev stage[kMaxStages];
  int n_ext; ExtDef ext[kMaxExt];
  // build sink
  LookupDev target; int bkey_col, target_unique, n_bpay, bpay_src[kMaxBuildPay], bpay_shift[kMaxBuildPay], bpay_width[kMaxBuildPay];
  // aggregate sink (group id == record of stage `agg_stage`)
  int agg_stage, rows_word, n_aggs; AggDef agg[kMaxPipeAggs];
  // unordered output sink
  int n_out, out_src[kMaxPipeCols], out_width[kMaxPipeCols]; void* out_dst[kMaxPipeCols]; unsigned long long* out_counter;
  ENode pool[kPoolNodes];
};

// ---- cache-policy loads: the table scan is read-once (evict first), the lookup structures should stay in L2 ----
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t ld_stream_int(const void* base, int width, int sgn, int64_t row, uint64_t pol) {
  switch (width) {
    case 1: { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"((const uint8_t*)base + row), "l"(pol)); return sgn ? (uint64_t)(int64_t)(int8_t)v : (uint64_t)v; }
    case 2: { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=r"(v) : "l"((const uint16_t*)base + row), "l"(pol)); return sgn ? (uint64_t)(int64_t)(int16_t)v : (uint64_t)v; }
    case 4: { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"((const uint32_t*)base + row), "l"(pol)); return sgn ? (uint64_t)(int64_t)(int32_t)v : (uint64_t)v; }
    default: { uint64_t v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"((const uint64_t*)base + row), "l"(pol)); return v; }
  }
}
__device__ __forceinline__ unsigned long long ld_keep_u64(const unsigned long long* p, uint64_t pol) {
  unsigned long long v; asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
}

struct alignas(16) Rec128 { unsigned long long lo, hi; };
__device__ __forceinline__ Rec128 rec_cas128(void* addr, Rec128 cmp, Rec128 val) {
  Rec128 old;
  asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\tatom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
               : "=l"(old.lo), "=l"(old.hi) : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr) : "memory");
  return old;
}

__device__ __forceinline__ uint64_t lk_hash(uint64_t key) { return hash_u64(key, kSeedJoin); }
// Coarse first level for filters that do not fit L2 (a join's filter shared by 4-8 GPUs is hundreds of MB): 4 bits per key, two probe
// bits in one 32-bit word.  It stays L2-resident and rejects ~85 % of the keys without a partner, so only ~1 probe in 4 pays the DRAM
// access of the exact (16 bits per key) level behind it.
struct CoarsePos { uint32_t word, mask; };
__device__ __forceinline__ CoarsePos coarse_pos(uint64_t key, uint64_t words) {
  uint32_t h = ((uint32_t)key * 0x27D4EB2Fu) ^ ((uint32_t)(key >> 32) * 0x165667B1u);
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13;
  CoarsePos p;
  p.word = __umulhi(h, (uint32_t)words);
  const uint32_t t = h * 0xC2B2AE35u;
  p.mask = (1u << (t >> 27)) | (1u << ((t >> 22) & 31));
  return p;
}
__device__ __forceinline__ void filter_set(const LookupDev& t, uint64_t key) {
  bloom_set(t.bloom, t.bloom_blocks, key);
  if (t.coarse) { const CoarsePos c = coarse_pos(key, t.coarse_words); atomicOr(&t.coarse[c.word], c.mask); }
}

// insert one record; returns 0 inserted, 1 duplicate key, 2 cannot store this key
__device__ __forceinline__ int lk_insert(const LookupDev& t, uint64_t key, uint64_t pay) {
  if (t.mode == LK_BITMAP) {
    const uint64_t i = key - t.kmin;
    if (i >= t.ksize) return 2;          // outside the promised range
    atomicOr(&t.bits[i >> 5], 1u << (i & 31));
    return 0;
  }
  if (key == kEmptyKey) return 2;
  const uint64_t h = lk_hash(key);
  if (t.cap == 0) {   // filter-only lookup: membership bits, no table
    filter_set(t, key);
    return 0;
  }
  uint64_t s = __umul64hi(h, t.cap);
  int rc = 0;
  while (true) {
    unsigned long long* r = t.recs + s * (uint64_t)t.stride;
    unsigned long long prev;
    if (t.has_payload) prev = rec_cas128(r, Rec128{kEmptyKey, 0ull}, Rec128{key, pay}).lo;
    else prev = atomicCAS(r, (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey) break;
    if (prev == key) { rc = 1; break; }
    if (++s == t.cap) s = 0;
  }
  if (rc == 0 && t.bloom) filter_set(t, key);
  return rc;
}

__device__ __forceinline__ uint64_t ext_field(uint64_t word, int shift, int width, int type) {
  uint64_t v = word >> shift;
  if (width < 8) { v &= (1ull << (8 * width)) - 1ull; if (type_is_signed_int(type)) v = (uint64_t)(((int64_t)(v << (64 - 8 * width))) >> (64 - 8 * width)); }
  return v;
}

__device__ __forceinline__ void red_f64_min(unsigned long long* p, double v, bool is_max) {
  unsigned long long old = *(volatile unsigned long long*)p;
  while (true) {
    const double cur = __longlong_as_double((long long)old);
    // f64 MIN/MAX follow the reference's total order on non-NaN data; NaN handling stays with the generic operator
    if (is_max ? !(v > cur) : !(v < cur)) return;
    const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) return;
    old = prev;
  }
}

// ------------------------------------------------------------------------------------------
// the pipeline kernel.  Every WARP runs the pipeline on its own 256-row tiles, in two phases, with no block barrier:
//   phase A (every row, cheap, coalesced): a lane owns 8 consecutive rows and reads them with 128-bit loads (a warp request
//     is 512 contiguous bytes per instruction); it evaluates the predicate, the bitmap stages and the Bloom pre-test of the
//     hash stages, and appends the surviving row numbers to the warp's queue in shared memory;
//   phase B (survivors only, dense): whenever the queue holds >= 128 entries every lane takes four of them — all lanes busy,
//     four table lookups in flight per lane — resolves the hash stages and feeds the sink (record insert / accumulator RED).
// Without the queue the expensive tail (a DRAM lookup, the argument interpreter, an insert CAS) runs at warp granularity for
// the few live lanes of every warp: that cost 27 warp-instructions and 143 B of DRAM traffic per row on Q3's lineitem pass.
// ------------------------------------------------------------------------------------------
constexpr int kWarpRows = 8, kWarpTile = 32 * kWarpRows, kPhaseB = 2, kPhaseBGroup = 32 * kPhaseB, kQueueCap = kPhaseBGroup + kWarpTile;
constexpr int kPipeWarps = kPipeThreads / 32;

__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_add_f64(unsigned long long* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void red_min_s64(unsigned long long* p, long long v) { asm volatile("red.global.min.s64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_max_s64(unsigned long long* p, long long v) { asm volatile("red.global.max.s64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_min_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.min.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_max_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.max.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// Pure 64-bit integer arithmetic (integer columns without NULLs in this batch, integer literals, payload fields, + - *): nothing can be
// NULL, nothing can fail — a four-register stack and ~10 instructions per node instead of the general interpreter's ~100.
__device__ __forceinline__ uint64_t eval_int_fast(const ENode* __restrict__ nodes, int n, int64_t row, const uint64_t* ext) {
  uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;   // s0 = top of stack
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const ENode& nd = nodes[i];
    if (nd.kind == DFGPU_EXPR_BINARY) {
      uint64_t r = nd.op == DFGPU_OP_PLUS ? s1 + s0 : (nd.op == DFGPU_OP_MINUS ? s1 - s0 : s1 * s0);
      if (type_width_prim(nd.out_type) < 8) r = wrap_to_type(r, nd.out_type);
      s0 = r; s1 = s2; s2 = s3;
    } else {
      uint64_t v;
      if (nd.kind == DFGPU_EXPR_COLUMN) v = load_col_value(nd, row);
      else if (nd.kind == DFGPU_EXPR_LITERAL) v = nd.lit;
      else { v = ext[nd.voff] >> (int)nd.lit; const int w = type_width_prim(nd.out_type); if (w < 8) { v &= (1ull << (8 * w)) - 1ull; if (type_is_signed_int(nd.out_type)) v = (uint64_t)(((int64_t)(v << (64 - 8 * w))) >> (64 - 8 * w)); } }
      s3 = s2; s2 = s1; s1 = s0; s0 = v;
    }
  }
  return s0;
}

// programs that touch Decimal128 values (small == 3): the 128-bit interpreter; returns the low word, *hi the high word
__device__ __noinline__ uint64_t pipe_eval_dec(const ENode* nodes, int n, int64_t row, const uint64_t* ext, int* err_ok, unsigned long long* hi) {
  bool ok;
  int err = 0;
  const i128 v = eval_nodes_dec(nodes, n, row, &ok, &err, ext);
  err_ok[0] |= err; err_ok[1] = ok ? 1 : 0;
  *hi = (unsigned long long)((u128)v >> 64);
  return (uint64_t)v;
}
// one out-of-line copy of each interpreter: the kernel stays small enough for the instruction cache
template <bool DEC>
__device__ __noinline__ uint64_t pipe_eval(const ENode* nodes, int n, int small, int64_t row, const uint64_t* ext, int* err_ok /* [0]=err bits (or-ed), [1]=valid */) {
  if (DEC && small == 3) { unsigned long long hi; return pipe_eval_dec(nodes, n, row, ext, err_ok, &hi); }
  bool ok;
  int err = 0;
  const uint64_t v = small ? eval_nodes_reg<4>(nodes, n, row, &ok, &err, ext) : eval_nodes(nodes, n, row, &ok, &err, ext);
  err_ok[0] |= err; err_ok[1] = ok ? 1 : 0;
  return v;
}
// SUM over a Decimal128 argument, evaluated and accumulated out of line (the hot integer path of the aggregate sink stays as it was):
// i128 add_wrapping over two accumulator words — the carry out of the low word is decided by this add alone
__device__ __noinline__ void pipe_sum_dec(const ENode* nodes, int n, int64_t row, const uint64_t* ext, int* err_ok, unsigned long long* acc, unsigned long long* nn) {
  unsigned long long hi;
  const unsigned long long lo = pipe_eval_dec(nodes, n, row, ext, err_ok, &hi);
  if (!err_ok[1]) return;                                  // NULL inputs are skipped (accumulate.rs:373-470)
  if (nn) atomicAdd(nn, 1ull);
  const unsigned long long old = atomicAdd(acc, lo);
  const unsigned long long add_hi = hi + ((old + lo) < old ? 1ull : 0ull);
  if (add_hi) atomicAdd(acc + 1, add_hi);
}

__device__ __forceinline__ uint4 ld_stream_v4(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_v2(const void* p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint64_t ext32(uint32_t x, int sgn) { return sgn ? (uint64_t)(int64_t)(int32_t)x : (uint64_t)x; }
__device__ __forceinline__ uint64_t ext16(uint32_t x, int sgn) { return sgn ? (uint64_t)(int64_t)(int16_t)x : (uint64_t)(x & 0xFFFFu); }
__device__ __forceinline__ uint64_t ext8(uint32_t x, int sgn) { return sgn ? (uint64_t)(int64_t)(int8_t)x : (uint64_t)(x & 0xFFu); }
// 8 consecutive elements starting at row0 (a multiple of 8), sign / zero extended to 64 bits
// 256-bit loads (LDG.E.ENL2.256, sm_100): a lane's 8 rows of a 4-byte column are ONE 32-byte sector, of an 8-byte column two — with
// 128-bit loads every instruction of the warp asks L2 for 32 HALF sectors (twice the requests for the same bytes)
__device__ __forceinline__ void ld_stream_v4x64(const void* p, uint64_t pol, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
}
template <bool WIDE = false>
__device__ __forceinline__ void load8(const ColRef& c, int64_t row0, int64_t n, uint64_t v[kWarpRows], uint64_t pol) {
  if (WIDE && c.vec == 2 && row0 + kWarpRows <= n && (c.width == 8 || c.width == 4)) {   // vec == 2: base pointer 32-byte aligned
    if (c.width == 8) {
      const char* p = (const char*)c.ptr + row0 * 8;
      ld_stream_v4x64(p, pol, v[0], v[1], v[2], v[3]);
      ld_stream_v4x64(p + 32, pol, v[4], v[5], v[6], v[7]);
    } else {
      uint64_t a, b, cc, d;
      ld_stream_v4x64((const char*)c.ptr + row0 * 4, pol, a, b, cc, d);
      v[0] = ext32((uint32_t)a, c.sgn); v[1] = ext32((uint32_t)(a >> 32), c.sgn); v[2] = ext32((uint32_t)b, c.sgn); v[3] = ext32((uint32_t)(b >> 32), c.sgn);
      v[4] = ext32((uint32_t)cc, c.sgn); v[5] = ext32((uint32_t)(cc >> 32), c.sgn); v[6] = ext32((uint32_t)d, c.sgn); v[7] = ext32((uint32_t)(d >> 32), c.sgn);
    }
    return;
  }
  if (c.vec && row0 + kWarpRows <= n) {
    switch (c.width) {
      case 8: {
        const char* p = (const char*)c.ptr + row0 * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint4 x = ld_stream_v4(p + 16 * q, pol); v[2 * q] = (uint64_t)x.x | ((uint64_t)x.y << 32); v[2 * q + 1] = (uint64_t)x.z | ((uint64_t)x.w << 32); }
        break;
      }
      case 4: {
        const char* p = (const char*)c.ptr + row0 * 4;
        const uint4 x = ld_stream_v4(p, pol), y = ld_stream_v4(p + 16, pol);
        v[0] = ext32(x.x, c.sgn); v[1] = ext32(x.y, c.sgn); v[2] = ext32(x.z, c.sgn); v[3] = ext32(x.w, c.sgn);
        v[4] = ext32(y.x, c.sgn); v[5] = ext32(y.y, c.sgn); v[6] = ext32(y.z, c.sgn); v[7] = ext32(y.w, c.sgn);
        break;
      }
      case 2: {
        const uint4 x = ld_stream_v4((const char*)c.ptr + row0 * 2, pol);
        v[0] = ext16(x.x, c.sgn); v[1] = ext16(x.x >> 16, c.sgn); v[2] = ext16(x.y, c.sgn); v[3] = ext16(x.y >> 16, c.sgn);
        v[4] = ext16(x.z, c.sgn); v[5] = ext16(x.z >> 16, c.sgn); v[6] = ext16(x.w, c.sgn); v[7] = ext16(x.w >> 16, c.sgn);
        break;
      }
      default: {
        const uint2 x = ld_stream_v2((const char*)c.ptr + row0, pol);
        v[0] = ext8(x.x, c.sgn); v[1] = ext8(x.x >> 8, c.sgn); v[2] = ext8(x.x >> 16, c.sgn); v[3] = ext8(x.x >> 24, c.sgn);
        v[4] = ext8(x.y, c.sgn); v[5] = ext8(x.y >> 8, c.sgn); v[6] = ext8(x.y >> 16, c.sgn); v[7] = ext8(x.y >> 24, c.sgn);
        break;
      }
    }
  } else {   // unaligned base or ragged tail: one scalar load site, results routed by selects (keeps v[] in registers, the code small)
#pragma unroll 1
    for (int j = 0; j < kWarpRows; ++j) {
      const uint64_t x = row0 + j < n ? ld_stream_int(c.ptr, c.width, c.sgn, row0 + j, pol) : 0ull;
#pragma unroll
      for (int k = 0; k < kWarpRows; ++k) if (k == j) v[k] = x;
    }
  }
}
// validity bits of the same 8 rows (bit j = row0 + j is non-NULL)
__device__ __forceinline__ uint32_t valid8(const ColRef& c, int64_t row0, int64_t n) {
  if (!c.valid) return 0xFFu;
  const int nb = (int)min((int64_t)kWarpRows, n - row0);
  return nb > 0 ? load_bits32(c.valid, c.voff + row0, nb) : 0u;
}

// VAR (compile-time, so that an instantiation which has been measured stays byte for byte what it was while others are tried against it):
//   bit 0 (1)  at the start of phase B, prefetch.global.L2 the survivors' argument sectors (the integer evaluator would read them one
//              dependent DRAM access after the other);
//   bit 1 (2)  at the start of phase A, prefetch this tile's key column of the first stage (its load is issued only after the predicate's
//              column has arrived and been compared);
//   bit 3 (8)  lane-paired REDs in the aggregate sink: a record's row counter and sum share a sector, so the even lane adds its row's count
//              while its odd neighbour adds the same row's value (then the roles swap) — one reduction request per row instead of two;
//   bit 5 (32) 256-bit column loads in phase A (LDG.E.ENL2.256): a lane's 8 rows are one or two whole sectors; with 128-bit loads every
//              instruction of the warp asks L2 for 32 half sectors.
// DFGPU_PIPE_VAR selects the instantiation (aggregate sink; bits 1 and 5 also for the pack sink, bit 1 for the unordered-output sink).  0 is the round-2 kernel as
// first measured (lineitem pass of Q3 at SF100: 8.1 ms), 11 = 7.5 ms, 43 the default = 7.1 ms (profiles/README.md "L2 request count").  Measured
// and removed: four instead of two survivors per lane and phase-B round (13.9 ms); prefetching the table record and the argument sectors
// already when a row passes the membership filter in phase A (10.4 ms: the prefetches of five tiles queue up in front of the column stream).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
template <int SINK, bool DEC, int VAR = 0>
__global__ void __launch_bounds__(kPipeThreads, 3) pipe_kernel(const PipeParams* __restrict__ gp, int64_t n, unsigned long long* __restrict__ counters /* [alive, inserted, fail, err] */) {
  constexpr int PB = kPhaseB, PBG = kPhaseBGroup, QC = kQueueCap;
  __shared__ PipeParams sp;
  __shared__ uint32_t q_rows[kPipeWarps][QC];
  for (int i = threadIdx.x; i < (int)(sizeof(PipeParams) / 4); i += kPipeThreads) ((uint32_t*)&sp)[i] = ((const uint32_t*)gp)[i];
  __syncthreads();
  const uint64_t pol_stream = (sp.hints & 1) ? policy_evict_first() : policy_normal();
  const uint64_t pol_keep = (sp.hints & 2) ? policy_evict_last() : policy_normal();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint32_t* q_row = q_rows[wib];
  unsigned int alive_cnt = 0, ins_cnt = 0;
  int err_ok[2] = {0, 0};
  int fail = 0;
  unsigned int qn = 0;   // queue length (warp-uniform)
  const int64_t ntiles = (n + kWarpTile - 1) / kWarpTile;
  const int64_t gwarp = (int64_t)blockIdx.x * kPipeWarps + wib, nwarps = (int64_t)gridDim.x * kPipeWarps;
  for (int64_t tile = gwarp; tile < ntiles + nwarps; tile += nwarps) {   // one extra trip per warp drains its queue
    const bool draining = tile >= ntiles;
    if (!draining) {
      // =============================== phase A ===============================
      const int64_t row0 = tile * kWarpTile + (int64_t)lane * kWarpRows;
      uint32_t mask = row0 + kWarpRows <= n ? 0xFFu : (row0 < n ? (1u << (int)(n - row0)) - 1u : 0u);
      if ((VAR & 2) && sp.n_stages > 0 && row0 < n) {
        const ColRef& kc0 = sp.col[sp.stage[0].key_col];
        prefetch_l2((const char*)kc0.ptr + row0 * kc0.width);
      }
      if (sp.pred_mode == 1) {   // FilterExec, conjunction of `column <cmp> literal`
#pragma unroll 1
        for (int t = 0; t < sp.n_terms; ++t) {
          const ColRef c = sp.col[sp.term_col[t]];
          uint64_t v[kWarpRows];
          load8<(VAR & 32) != 0>(c, row0, n, v, pol_stream);
          mask &= valid8(c, row0, n);            // a NULL predicate drops the row
          const int op = sp.term_op[t];
          const long long lit = sp.term_lit[t];
          uint32_t lt = 0, eq = 0;
          if (sp.term_uns[t]) {
#pragma unroll
            for (int j = 0; j < kWarpRows; ++j) { lt |= (uint32_t)(v[j] < (uint64_t)lit) << j; eq |= (uint32_t)(v[j] == (uint64_t)lit) << j; }
          } else {
#pragma unroll
            for (int j = 0; j < kWarpRows; ++j) { lt |= (uint32_t)((long long)v[j] < lit) << j; eq |= (uint32_t)((long long)v[j] == lit) << j; }
          }
          uint32_t r;
          switch (op) {
            case DFGPU_OP_EQ: r = eq; break;
            case DFGPU_OP_NEQ: r = ~eq; break;
            case DFGPU_OP_LT: r = lt; break;
            case DFGPU_OP_LTEQ: r = lt | eq; break;
            case DFGPU_OP_GT: r = ~(lt | eq); break;
            default: r = ~lt; break;
          }
          mask &= r;
        }
      } else if (sp.pred_mode == 2) {   // FilterExec, general expression
#pragma unroll 1
        for (int j = 0; j < kWarpRows; ++j) {
          if (!((mask >> j) & 1u)) continue;
          const uint64_t val = pipe_eval<DEC>(sp.pool + sp.pred_start, sp.pred_n, sp.pred_small, row0 + j, nullptr, err_ok);
          if (!(err_ok[1] && (val & 1))) mask &= ~(1u << j);
        }
      }
      // bitmap stages decide here; hash stages get their Bloom pre-test (the pushed-down membership filter)
#pragma unroll 1
      for (int s = 0; s < sp.n_stages; ++s) {
        const StageDev& st = sp.stage[s];
        const bool bitmap = st.lk.mode == LK_BITMAP;
        if (!bitmap && !(st.lk.bloom && st.kind != DFGPU_STAGE_ANTI)) {   // nothing cheap to test; NULL keys of an inner / semi stage still drop here
          if (st.kind != DFGPU_STAGE_ANTI && sp.col[st.key_col].valid) mask &= valid8(sp.col[st.key_col], row0, n);
          continue;
        }
        if (!bitmap && st.kind == kStageMaybe && !st.lk.bloom) continue;                                // may-contain without a filter: everything may
        if (!__any_sync(0xffffffffu, mask != 0)) continue;
        const ColRef kc = sp.col[st.key_col];
        uint64_t key[kWarpRows];
        load8<(VAR & 32) != 0>(kc, row0, n, key, pol_stream);
        const uint32_t kvalid = valid8(kc, row0, n);
        if (bitmap) {
          uint32_t w[kWarpRows];
#pragma unroll
          for (int j = 0; j < kWarpRows; ++j) {
            const uint64_t i = key[j] - st.lk.kmin;
            w[j] = (((mask & kvalid) >> j) & 1u) && i < st.lk.ksize ? __ldg(&st.lk.bits[i >> 5]) : 0u;
          }
          uint32_t found = 0;
#pragma unroll
          for (int j = 0; j < kWarpRows; ++j) found |= ((w[j] >> ((key[j] - st.lk.kmin) & 31)) & 1u) << j;   // NULL keys never match (w = 0)
          mask &= st.kind == DFGPU_STAGE_ANTI ? ~found : found;
        } else {
          if (st.kind != DFGPU_STAGE_ANTI) {
            mask &= kvalid;                      // NULL keys never match
            if (st.lk.coarse) {   // first level: L2-resident, removes most rows before the exact level's DRAM access
              uint32_t cw[kWarpRows], cm[kWarpRows];
#pragma unroll
              for (int j = 0; j < kWarpRows; ++j) {
                const CoarsePos cp = coarse_pos(key[j], st.lk.coarse_words);
                cm[j] = cp.mask;
                cw[j] = ((mask >> j) & 1u) ? __ldg(&st.lk.coarse[cp.word]) : 0u;
              }
              uint32_t cpass = 0;
#pragma unroll
              for (int j = 0; j < kWarpRows; ++j) cpass |= (uint32_t)((cw[j] & cm[j]) == cm[j]) << j;
              mask &= cpass;
            }
            if (st.lk.bloom) {
              unsigned long long bw[kWarpRows];
              uint32_t bt[kWarpRows];
#pragma unroll
              for (int j = 0; j < kWarpRows; ++j) {
                const BloomPos bp = bloom_pos(key[j], st.lk.bloom_blocks);
                bt[j] = bp.t;
                bw[j] = ((mask >> j) & 1u) ? ld_keep_u64(&st.lk.bloom[bp.block], pol_keep) : 0ull;
              }
              uint32_t pass = 0;
#pragma unroll
              for (int j = 0; j < kWarpRows; ++j) pass |= (uint32_t)bloom_test(bw[j], bt[j]) << j;
              mask &= pass;
            }
          }
        }
      }
      // append the survivors to the warp's queue (exclusive scan of the per-lane counts)
      const unsigned int cnt = __popc(mask);
      unsigned int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      unsigned int pos = qn + incl - cnt;
#pragma unroll
      for (int j = 0; j < kWarpRows; ++j)
        if ((mask >> j) & 1u) { q_row[pos] = (uint32_t)(row0 + j); ++pos; }
      qn += __shfl_sync(0xffffffffu, incl, 31);
      __syncwarp();
    }
    // =============================== phase B ===============================
    while (qn >= (unsigned)PBG || (draining && qn > 0)) {
      const unsigned int take = qn >= (unsigned)PBG ? (unsigned)PBG : qn;
      const unsigned int qbase = qn - take;   // consume from the tail: nothing has to move
      bool live[PB];
      int64_t row[PB];
      uint64_t pay[kMaxStages][PB];
      unsigned long long* arec[PB];
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const unsigned int e = u * 32 + lane;
        live[u] = e < take; row[u] = live[u] ? (int64_t)q_row[qbase + e] : 0; arec[u] = nullptr;
      }
      if ((VAR & 1) && SINK == SINK_AGG) {
#pragma unroll 1
        for (int a = 0; a < sp.n_aggs; ++a) {
          const AggDef& ag = sp.agg[a];
          if (ag.small != 2) continue;
#pragma unroll 1
          for (int i = 0; i < ag.n; ++i) {
            const ENode& nd = sp.pool[ag.start + i];
            if (nd.kind != DFGPU_EXPR_COLUMN) continue;
            const int w = type_width_prim(nd.out_type);
#pragma unroll
            for (int u = 0; u < PB; ++u) if (live[u]) prefetch_l2((const char*)nd.col + row[u] * w);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) {
#pragma unroll
        for (int u = 0; u < PB; ++u) pay[s][u] = 0;
        if (s >= sp.n_stages) continue;
        const StageDev& st = sp.stage[s];
        if (st.lk.mode == LK_BITMAP || st.kind == kStageMaybe) continue;   // decided in phase A
        const ColRef kc = sp.col[st.key_col];
        uint64_t key[PB], slot[PB], ck[PB], cp[PB];
        bool look[PB], found[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          found[u] = false; look[u] = live[u]; key[u] = 0;
          if (look[u]) {
            key[u] = ld_stream_int(kc.ptr, kc.width, kc.sgn, row[u], pol_stream);   // the line was streamed in moments ago: an L2 hit
            if ((kc.valid && !bit_get(kc.valid, kc.voff + row[u])) || key[u] == kEmptyKey) look[u] = false;   // NULL keys never match
          }
          slot[u] = __umul64hi(lk_hash(key[u]), st.lk.cap); ck[u] = kEmptyKey; cp[u] = 0;
          if (look[u]) {
            const unsigned long long* r = st.lk.recs + slot[u] * (uint64_t)st.lk.stride;
            if (st.lk.has_payload) { const uint4 v = __ldcg((const uint4*)r); ck[u] = (uint64_t)v.x | ((uint64_t)v.y << 32); cp[u] = (uint64_t)v.z | ((uint64_t)v.w << 32); }
            else ck[u] = __ldcg(r);
          }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          if (look[u]) {
            while (true) {   // linear probing continues only past a foreign key (load factor <= 0.5)
              if (ck[u] == key[u]) { found[u] = true; break; }
              if (ck[u] == kEmptyKey) break;
              if (++slot[u] == st.lk.cap) slot[u] = 0;
              const unsigned long long* r = st.lk.recs + slot[u] * (uint64_t)st.lk.stride;
              if (st.lk.has_payload) { const uint4 v = __ldcg((const uint4*)r); ck[u] = (uint64_t)v.x | ((uint64_t)v.y << 32); cp[u] = (uint64_t)v.z | ((uint64_t)v.w << 32); }
              else ck[u] = __ldcg(r);
            }
            if (found[u]) { pay[s][u] = cp[u]; if (SINK == SINK_AGG && s == sp.agg_stage) arec[u] = st.lk.recs + slot[u] * (uint64_t)st.lk.stride; }
          }
          live[u] = live[u] && (st.kind == DFGPU_STAGE_ANTI ? !found[u] : found[u]);
        }
      }
      // ---- sink ----
      if (SINK == SINK_OUTPUT_ANY) {   // rows leave in arrival order: one global reservation per 128-row group, coalesced column writes
        unsigned int tot = 0, mypos[PB];
#pragma unroll
        for (int u = 0; u < PB; ++u) { const unsigned m = __ballot_sync(0xffffffffu, live[u]); mypos[u] = tot + __popc(m & ((1u << lane) - 1u)); tot += __popc(m); }
        unsigned long long obase = 0;
        if (lane == 0 && tot) obase = atomicAdd(sp.out_counter, (unsigned long long)tot);
        obase = __shfl_sync(0xffffffffu, obase, 0);
#pragma unroll
        for (int u = 0; u < PB; ++u) {
          if (!live[u]) continue;
          alive_cnt++;
          for (int c = 0; c < sp.n_out; ++c) {
            const int src = sp.out_src[c], w = sp.out_width[c];
            uint64_t v;
            if (src < sp.n_cols) v = ld_stream_int(sp.col[src].ptr, w, 0, row[u], pol_stream);
            else { const ExtDef e = sp.ext[src - sp.n_cols]; uint64_t wd = 0;
#pragma unroll
                   for (int s = 0; s < kMaxStages; ++s) if (s == e.stage) wd = pay[s][u];
                   v = ext_field(wd, e.shift, e.width, DFGPU_UINT64); }
            const unsigned long long o = obase + mypos[u];
            switch (w) {
              case 1: ((uint8_t*)sp.out_dst[c])[o] = (uint8_t)v; break;
              case 2: ((uint16_t*)sp.out_dst[c])[o] = (uint16_t)v; break;
              case 4: ((uint32_t*)sp.out_dst[c])[o] = (uint32_t)v; break;
              default: ((uint64_t*)sp.out_dst[c])[o] = v; break;
            }
          }
        }
      }
      bool sunk = false;
      if ((VAR & 8) && SINK == SINK_AGG) {
        // one RED instruction per lane PAIR and row: the row counter and the sum word of a record share a sector, so the even lane adds
        // its row's count while the odd neighbour adds the same row's value (then the roles swap) — half the L2 reduction requests
        const AggDef& ag0 = sp.agg[0];
        if (sp.n_aggs == 1 && ag0.small == 2 && ag0.func == DFGPU_AGG_SUM && ag0.cls != C_F64 && ag0.cls != C_DEC && ag0.nn_word < 0) {
          sunk = true;
          const bool even = !(lane & 1);
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            uint64_t ext[kMaxStages];
#pragma unroll
            for (int s = 0; s < kMaxStages; ++s) ext[s] = pay[s][u];
            unsigned long long v = 0;
            if (live[u]) { alive_cnt++; v = eval_int_fast(sp.pool + ag0.start, ag0.n, row[u], ext); }
            __syncwarp();
            const unsigned long long rp = (unsigned long long)arec[u];
            const unsigned long long prp = __shfl_xor_sync(0xffffffffu, rp, 1);
            const unsigned long long pv = __shfl_xor_sync(0xffffffffu, v, 1);
            const bool pl = __shfl_xor_sync(0xffffffffu, live[u] ? 1 : 0, 1) != 0;
            unsigned long long* const mine = (unsigned long long*)rp + sp.rows_word;    // this lane's row: the row counter
            unsigned long long* const theirs = (unsigned long long*)prp + ag0.word;     // the neighbour's row: the sum
            if (even ? live[u] : pl) red_add_u64(even ? mine : theirs, even ? 1ull : pv);     // rows of the even lanes
            if (even ? pl : live[u]) red_add_u64(even ? theirs : mine, even ? pv : 1ull);     // rows of the odd lanes
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        if (SINK == SINK_OUTPUT_ANY || sunk) break;
        uint64_t ext[kMaxStages];
#pragma unroll
        for (int s = 0; s < kMaxStages; ++s) ext[s] = pay[s][u];
        if (!live[u]) continue;
        alive_cnt++;
        if (SINK == SINK_BUILD || SINK == SINK_PACK) {
          const ColRef kc = sp.col[sp.bkey_col];
          if (kc.valid && !bit_get(kc.valid, kc.voff + row[u])) continue;   // NULL build keys are not inserted (utils.rs:2146-2155)
          const uint64_t key = ld_stream_int(kc.ptr, kc.width, kc.sgn, row[u], pol_stream);
          uint64_t p = 0;
          for (int c = 0; c < sp.n_bpay; ++c) {
            const int src = sp.bpay_src[c];
            uint64_t v;
            if (src < sp.n_cols) v = ld_stream_int(sp.col[src].ptr, sp.col[src].width, 0, row[u], pol_stream);
            else { const ExtDef e = sp.ext[src - sp.n_cols]; v = ext_field(ext[e.stage], e.shift, e.width, DFGPU_UINT64); }
            if (sp.bpay_width[c] < 8) v &= (1ull << (8 * sp.bpay_width[c])) - 1ull;
            p |= v << sp.bpay_shift[c];
          }
          if (SINK == SINK_PACK) {   // one reservation per warp and round (the lanes still active here), coalesced 16-byte stores
            const unsigned am = __activemask();
            const int leader = __ffs(am) - 1;
            unsigned long long o = 0;
            if (lane == leader) o = atomicAdd(sp.out_counter, (unsigned long long)__popc(am));
            o = __shfl_sync(am, o, leader) + __popc(am & ((1u << lane) - 1u));
            ((ulonglong2*)sp.out_dst[0])[o] = make_ulonglong2(key, p);
            ins_cnt++;
          } else {
            const int rc = lk_insert(sp.target, key, p);
            if (rc == 0) ins_cnt++;
            else if (rc == 2 || sp.target_unique) fail |= rc;
          }
        } else if (SINK == SINK_AGG) {
          unsigned long long* rec = arec[u];
          red_add_u64(rec + sp.rows_word, 1ull);
          for (int a = 0; a < sp.n_aggs; ++a) {
            const AggDef ag = sp.agg[a];
            if (ag.func == DFGPU_AGG_COUNT_STAR) continue;   // = the row counter
            uint64_t v;
            if (ag.small == 2) v = eval_int_fast(sp.pool + ag.start, ag.n, row[u], ext);
            else if (DEC && ag.cls == C_DEC && ag.func == DFGPU_AGG_SUM) {
              pipe_sum_dec(sp.pool + ag.start, ag.n, row[u], ext, err_ok, rec + ag.word, ag.nn_word >= 0 ? rec + ag.nn_word : nullptr);
              continue;
            } else {
              v = pipe_eval<DEC>(sp.pool + ag.start, ag.n, ag.small, row[u], ext, err_ok);
              if (!err_ok[1]) continue;                      // NULL inputs are skipped (accumulate.rs:373-470)
            }
            if (ag.nn_word >= 0) red_add_u64(rec + ag.nn_word, 1ull);
            switch (ag.func) {
              case DFGPU_AGG_COUNT: red_add_u64(rec + ag.word, 1ull); break;
              case DFGPU_AGG_SUM: case DFGPU_AGG_AVG:
                if (ag.cls == C_F64) red_add_f64(rec + ag.word, __longlong_as_double((long long)v));
                else red_add_u64(rec + ag.word, (unsigned long long)v);   // add_wrapping (sum.rs:316)
                break;
              case DFGPU_AGG_MIN:
                if (ag.cls == C_F64) red_f64_min(rec + ag.word, __longlong_as_double((long long)v), false);
                else if (ag.cls == C_U64) red_min_u64(rec + ag.word, (unsigned long long)v);
                else red_min_s64(rec + ag.word, (long long)v);
                break;
              case DFGPU_AGG_MAX:
                if (ag.cls == C_F64) red_f64_min(rec + ag.word, __longlong_as_double((long long)v), true);
                else if (ag.cls == C_U64) red_max_u64(rec + ag.word, (unsigned long long)v);
                else red_max_s64(rec + ag.word, (long long)v);
                break;
            }
          }
        }
      }
      qn = qbase;
      __syncwarp();
    }
  }
  // block-level counter reduction: one atomic per block and counter
  __shared__ unsigned int s_red[2][kPipeWarps];
  __shared__ int s_flag[2];
  if (threadIdx.x == 0) { s_flag[0] = 0; s_flag[1] = 0; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { alive_cnt += __shfl_xor_sync(0xffffffffu, alive_cnt, d); ins_cnt += __shfl_xor_sync(0xffffffffu, ins_cnt, d); }
  __syncthreads();
  if (lane == 0) { s_red[0][wib] = alive_cnt; s_red[1][wib] = ins_cnt; }
  if (fail) atomicOr(&s_flag[0], fail);
  if (err_ok[0]) atomicOr(&s_flag[1], err_ok[0]);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int w = 0; w < kPipeWarps; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
    if (a) atomicAdd(&counters[0], a);
    if (b) atomicAdd(&counters[1], b);
    if (s_flag[0]) atomicOr(&counters[2], (unsigned long long)s_flag[0]);
    if (s_flag[1]) atomicOr(&counters[3], (unsigned long long)s_flag[1]);
  }
}

// ------------------------------------------------------------------------------------------
// output sink: surviving rows in input order (ordered compaction: block scan + decoupled look-back, as filter_fused_kernel)
// ------------------------------------------------------------------------------------------
struct OutCols { int n; int src[kMaxPipeCols]; int width[kMaxPipeCols]; void* dst[kMaxPipeCols]; };

__global__ void __launch_bounds__(kPipeThreads) pipe_output_kernel(const PipeParams* __restrict__ gp, int64_t n, OutCols oc, unsigned long long* __restrict__ tile_desc,
                                                                  unsigned int* __restrict__ tile_counter, unsigned long long* __restrict__ totals, unsigned long long* __restrict__ counters) {
  __shared__ PipeParams sp;
  __shared__ uint32_t s_p[kPipeTile];
  __shared__ unsigned long long s_pay[kMaxStages][kPipeTile];
  __shared__ unsigned int s_tile;
  __shared__ unsigned long long s_base;
  for (int i = threadIdx.x; i < (int)(sizeof(PipeParams) / 4); i += kPipeThreads) ((uint32_t*)&sp)[i] = ((const uint32_t*)gp)[i];
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t row0 = tile * kPipeTile + (int64_t)threadIdx.x * kPipeItems;   // consecutive rows per thread: rank order == row order
  int err = 0;
  bool alive[kPipeItems];
  uint64_t pay[kMaxStages][kPipeItems];
#pragma unroll
  for (int k = 0; k < kPipeItems; ++k) {
    const int64_t row = row0 + k;
    alive[k] = row < n;
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) pay[s][k] = 0;
    if (!alive[k]) continue;
    if (sp.pred_mode == 1) {
      for (int t = 0; t < sp.n_terms && alive[k]; ++t) {
        const ColRef c = sp.col[sp.term_col[t]];
        if (c.valid && !bit_get(c.valid, c.voff + row)) { alive[k] = false; break; }
        const uint64_t v = ld_stream_int(c.ptr, c.width, c.sgn, row, policy_evict_first());
        const long long lit = sp.term_lit[t];
        int cmp;
        if (sp.term_uns[t]) cmp = v < (uint64_t)lit ? -1 : (v > (uint64_t)lit ? 1 : 0);
        else cmp = (long long)v < lit ? -1 : ((long long)v > lit ? 1 : 0);
        switch (sp.term_op[t]) {
          case DFGPU_OP_EQ: alive[k] = cmp == 0; break;
          case DFGPU_OP_NEQ: alive[k] = cmp != 0; break;
          case DFGPU_OP_LT: alive[k] = cmp < 0; break;
          case DFGPU_OP_LTEQ: alive[k] = cmp <= 0; break;
          case DFGPU_OP_GT: alive[k] = cmp > 0; break;
          default: alive[k] = cmp >= 0; break;
        }
      }
    } else if (sp.pred_mode == 2) {
      bool ok;
      const uint64_t val = eval_nodes(sp.pool + sp.pred_start, sp.pred_n, row, &ok, &err);
      alive[k] = ok && (val & 1);
    }
  }
#pragma unroll
  for (int s = 0; s < kMaxStages; ++s) {
    if (s >= sp.n_stages) continue;
    const StageDev& st = sp.stage[s];
    const ColRef kc = sp.col[st.key_col];
#pragma unroll
    for (int k = 0; k < kPipeItems; ++k) {
      if (!alive[k]) continue;
      const int64_t row = row0 + k;
      bool found = false;
      const bool knull = kc.valid && !bit_get(kc.valid, kc.voff + row);
      const uint64_t key = ld_stream_int(kc.ptr, kc.width, kc.sgn, row, policy_evict_first());
      if (!knull) {
        if (st.lk.mode == LK_BITMAP) {
          const uint64_t i = key - st.lk.kmin;
          found = i < st.lk.ksize && ((__ldg(&st.lk.bits[i >> 5]) >> (i & 31)) & 1u);
        } else if (key != kEmptyKey) {
          const uint64_t h = lk_hash(key);
          bool maybe = true;
          if (st.lk.coarse) { const CoarsePos cp = coarse_pos(key, st.lk.coarse_words); maybe = (st.lk.coarse[cp.word] & cp.mask) == cp.mask; }
          if (maybe && st.lk.bloom) { const BloomPos bp = bloom_pos(key, st.lk.bloom_blocks); maybe = bloom_test(st.lk.bloom[bp.block], bp.t); }
          if (maybe) {
            uint64_t slot = __umul64hi(h, st.lk.cap);
            while (true) {
              const unsigned long long* r = st.lk.recs + slot * (uint64_t)st.lk.stride;
              const unsigned long long ck = __ldcg(r);
              if (ck == key) { found = true; if (st.lk.has_payload) pay[s][k] = __ldcg(r + 1); break; }
              if (ck == kEmptyKey) break;
              if (++slot == st.lk.cap) slot = 0;
            }
          }
        }
      }
      alive[k] = st.kind == DFGPU_STAGE_ANTI ? !found : found;
    }
  }
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < kPipeItems; ++k) m += alive[k] ? 1u : 0u;
  uint32_t tot;
  uint32_t ex = block_exclusive_scan<kPipeThreads, uint32_t>(m, &tot);
#pragma unroll
  for (int k = 0; k < kPipeItems; ++k)
    if (alive[k]) {
      s_p[ex] = (uint32_t)(threadIdx.x * kPipeItems + k);
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) s_pay[s][ex] = pay[s][k];
      ++ex;
    }
  if (threadIdx.x < 32) {
    unsigned long long exclusive = tile_lookback(tile, tot, tile_desc);
    if (threadIdx.x == 0) {
      s_base = exclusive;
      if ((tile + 1) * (int64_t)kPipeTile >= n) totals[0] = exclusive + tot;
    }
  }
  __syncthreads();
  const unsigned long long obase = s_base;
  const int64_t prow0 = tile * kPipeTile;
  for (int c = 0; c < oc.n; ++c) {
    const int src = oc.src[c], w = oc.width[c];
    for (uint32_t j = threadIdx.x; j < tot; j += kPipeThreads) {
      uint64_t v;
      if (src < sp.n_cols) v = ld_stream_int(sp.col[src].ptr, w, 0, prow0 + s_p[j], policy_evict_first());
      else { const ExtDef e = sp.ext[src - sp.n_cols]; v = ext_field(s_pay[e.stage][j], e.shift, e.width, DFGPU_UINT64); }
      switch (w) {
        case 1: ((uint8_t*)oc.dst[c])[obase + j] = (uint8_t)v; break;
        case 2: ((uint16_t*)oc.dst[c])[obase + j] = (uint16_t)v; break;
        case 4: ((uint32_t*)oc.dst[c])[obase + j] = (uint32_t)v; break;
        default: ((uint64_t*)oc.dst[c])[obase + j] = v; break;
      }
    }
  }
  if (err) atomicOr(&counters[3], (unsigned long long)err);
}

// ------------------------------------------------------------------------------------------
// lookup maintenance kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lookup_init_kernel(unsigned long long* recs, uint64_t cap, int stride) {
  const uint64_t total = cap * (uint64_t)stride;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    recs[i] = (i % (uint64_t)stride) == 0 ? kEmptyKey : 0ull;
}
__global__ void __launch_bounds__(256) lookup_rehash_kernel(LookupDev old_t, LookupDev new_t) {
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < old_t.cap; s += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long* r = old_t.recs + s * (uint64_t)old_t.stride;
    const unsigned long long key = r[0];
    if (key == kEmptyKey) continue;
    const uint64_t h = lk_hash(key);
    uint64_t d = __umul64hi(h, new_t.cap);
    while (true) {
      unsigned long long* q = new_t.recs + d * (uint64_t)new_t.stride;
      if (atomicCAS(q, (unsigned long long)kEmptyKey, key) == kEmptyKey) { for (int w = 1; w < new_t.stride; ++w) q[w] = r[w]; break; }
      if (++d == new_t.cap) d = 0;
    }
    if (new_t.bloom) filter_set(new_t, key);
  }
}
// second half of a build whose size was unknown: the packed {key, payload} records of the survivors go into the (now sized) table
__global__ void __launch_bounds__(256) lookup_insert_records_kernel(LookupDev t, const ulonglong2* __restrict__ recs, int64_t n, int unique, unsigned long long* __restrict__ counters /* [_, inserted, fail] */) {
  unsigned int ins = 0; int fail = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const ulonglong2 r = recs[i];
    const int rc = lk_insert(t, r.x, r.y);
    if (rc == 0) ins++;
    else if (rc == 2 || unique) fail |= rc;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ins += __shfl_xor_sync(0xffffffffu, ins, d);
  fail = __any_sync(0xffffffffu, fail & 1) | (__any_sync(0xffffffffu, fail & 2) << 1);
  if ((threadIdx.x & 31) == 0) { if (ins) atomicAdd(&counters[1], (unsigned long long)ins); if (fail) atomicOr(&counters[2], (unsigned long long)fail); }
}
// accumulator identities for MIN / MAX (SUM / COUNT start at the zero the table was initialised with)
__global__ void __launch_bounds__(256) lookup_init_acc_kernel(LookupDev t, int word, unsigned long long value) {
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < t.cap; s += (uint64_t)gridDim.x * blockDim.x) t.recs[s * (uint64_t)t.stride + word] = value;
}
// records with rows_word > 0 -> occupancy bitmap (one ballot word per warp)
__global__ void __launch_bounds__(256) lookup_groups_kernel(LookupDev t, int rows_word, uint32_t* __restrict__ words) {
  const uint64_t nw = (t.cap + 31) / 32;
  const int lane = threadIdx.x & 31;
  for (uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5; w < nw; w += ((uint64_t)gridDim.x * blockDim.x) >> 5) {
    const uint64_t s = w * 32 + lane;
    const bool occ = s < t.cap && t.recs[s * (uint64_t)t.stride + rows_word] != 0ull;
    const uint32_t b = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) words[w] = b;
  }
}
struct EmitCol { int kind /* 0 key, 1 payload field, 2 accumulator word, 3 AVG value, 4 count as u64, 5 Decimal128 sum (two words) */, width, shift, word, nn_word, cnt_word, f64; void* dst; uint32_t* valid; };
struct EmitCols { int n; EmitCol c[kMaxPipeCols]; };
__global__ void __launch_bounds__(256) lookup_emit_kernel(LookupDev t, const uint32_t* __restrict__ slots, int64_t n, int rows_word, EmitCols ec) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (n + 31) / 32;
  for (int64_t wi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; wi < nw; wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t i = wi * 32 + lane;
    const unsigned long long* r = i < n ? t.recs + (uint64_t)slots[i] * (uint64_t)t.stride : nullptr;
    for (int c = 0; c < ec.n; ++c) {
      const EmitCol e = ec.c[c];
      uint64_t v = 0;
      bool ok = false;
      if (r) {
        ok = true;
        switch (e.kind) {
          case 0: v = r[0]; break;
          case 1: v = r[1] >> e.shift; break;
          case 2: v = r[e.word]; ok = e.nn_word >= 0 ? r[e.nn_word] != 0ull : true; break;
          case 4: v = r[e.word]; break;
          case 5:
            ok = e.nn_word >= 0 ? r[e.nn_word] != 0ull : true;
            ((unsigned long long*)e.dst)[2 * i] = ok ? r[e.word] : 0ull;
            ((unsigned long long*)e.dst)[2 * i + 1] = ok ? r[e.word + 1] : 0ull;
            break;
          default: {   // AVG = sum / count over Float64 (functions-aggregate/src/average.rs)
            const unsigned long long cnt = r[e.cnt_word];
            ok = cnt != 0ull;
            const double d = ok ? __longlong_as_double((long long)r[e.word]) / (double)cnt : 0.0;
            v = (uint64_t)__double_as_longlong(d);
          }
        }
        if (!ok) v = 0;
        if (e.kind != 5) switch (e.width) {
          case 1: ((uint8_t*)e.dst)[i] = (uint8_t)v; break;
          case 2: ((uint16_t*)e.dst)[i] = (uint16_t)v; break;
          case 4: ((uint32_t*)e.dst)[i] = (uint32_t)v; break;
          default: ((uint64_t*)e.dst)[i] = v; break;
        }
      }
      if (e.valid) { const uint32_t b = __ballot_sync(0xffffffffu, ok); if (lane == 0) e.valid[wi] = b; }
    }
  }
}

// OR-all-reduce of n_ranks membership filters of identical geometry over peer memory (NVLink): this rank merges slice
// `rank` of every filter (reads of the peers' slices travel over NVLink) and writes the merged slice into every rank's
// filter.  Slice r of rank q's buffer is read only by rank r, and written by rank r only after it has read it.
constexpr int kMaxPeers = 8;
struct PeerWords { unsigned long long* p[kMaxPeers]; };
__global__ void __launch_bounds__(256) filter_allreduce_peer_kernel(PeerWords pw, int rank, int n_ranks, uint64_t blocks) {
  const uint64_t lo = blocks * (uint64_t)rank / (uint64_t)n_ranks, hi = blocks * (uint64_t)(rank + 1) / (uint64_t)n_ranks;
  for (uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long acc = 0;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) if (q < n_ranks) acc |= pw.p[q][i];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) if (q < n_ranks) pw.p[q][i] = acc;
  }
}

__global__ void __launch_bounds__(256) col_minmax_kernel(ColRef c, int64_t n, int uns, unsigned long long* mm /* [min,max,valid] */) {
  unsigned long long kmin = ~0ull, kmax = 0, cnt = 0;   // order-preserving map of signed keys onto unsigned
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (c.valid && !bit_get(c.valid, c.voff + i)) continue;
    uint64_t v;
    switch (c.width) {
      case 1: v = c.sgn ? (uint64_t)(int64_t)((const int8_t*)c.ptr)[i] : ((const uint8_t*)c.ptr)[i]; break;
      case 2: v = c.sgn ? (uint64_t)(int64_t)((const int16_t*)c.ptr)[i] : ((const uint16_t*)c.ptr)[i]; break;
      case 4: v = c.sgn ? (uint64_t)(int64_t)((const int32_t*)c.ptr)[i] : ((const uint32_t*)c.ptr)[i]; break;
      default: v = ((const uint64_t*)c.ptr)[i]; break;
    }
    if (!uns) v ^= 1ull << 63;
    kmin = min(kmin, (unsigned long long)v); kmax = max(kmax, (unsigned long long)v); cnt++;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, d));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0 && cnt) { atomicMin(&mm[0], kmin); atomicMax(&mm[1], kmax); atomicAdd(&mm[2], cnt); }
}

// wrapping sum of an integer column (sign / zero extended to 64 bits): order-independent fingerprints of large results
__global__ void __launch_bounds__(256) col_sum_kernel(ColRef c, int64_t n, unsigned long long* out /* [sum, valid] */) {
  unsigned long long s = 0, cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (c.valid && !bit_get(c.valid, c.voff + i)) continue;
    switch (c.width) {
      case 1: s += c.sgn ? (uint64_t)(int64_t)((const int8_t*)c.ptr)[i] : ((const uint8_t*)c.ptr)[i]; break;
      case 2: s += c.sgn ? (uint64_t)(int64_t)((const int16_t*)c.ptr)[i] : ((const uint16_t*)c.ptr)[i]; break;
      case 4: s += c.sgn ? (uint64_t)(int64_t)((const int32_t*)c.ptr)[i] : ((const uint32_t*)c.ptr)[i]; break;
      case 16: s += ((const uint64_t*)c.ptr)[2 * i] + 3ull * ((const uint64_t*)c.ptr)[2 * i + 1]; break;   // Decimal128: low word + 3 x high word
      default: s += ((const uint64_t*)c.ptr)[i]; break;
    }
    cnt++;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); cnt += __shfl_xor_sync(0xffffffffu, cnt, d); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], s); atomicAdd(&out[1], cnt); }
}

}  // namespace dfgpu

// ==========================================================================================
// host side
// ==========================================================================================
using namespace dfgpu;

struct dfgpu_lookup {
  dfgpu_ctx* ctx = nullptr;
  int key_type = 0;
  std::vector<int> pay_types, pay_shift;
  dfgpu_lookup_options opt{};
  int mode = LK_HASH, stride = 1;
  bool has_payload = false;
  DevBuf recs, bloom, bits;
  uint64_t cap = 0, bloom_blocks = 0, coarse_words = 0, kmin = 0, ksize = 0;   // the coarse level lives behind the exact blocks in `bloom`
  int64_t rows = 0, rehashes = 0;
  bool acc_claimed = false, filter_only = false;
};

struct PipeAgg { int func; ExprPlan plan; bool has_expr = false; int word = -1, nn_word = -1, cnt_word = -1, cls = C_I64, arg_type = 0; };

struct dfgpu_pipeline {
  dfgpu_ctx* ctx = nullptr;
  std::vector<int> in_types, vtypes;                 // input schema, virtual schema (input + payload fields)
  std::vector<ExtDef> exts;
  bool has_pred = false;
  ExprPlan pred;
  std::vector<dfgpu_pipeline_stage> stages;
  int sink = SINK_NONE;
  // build sink
  dfgpu_lookup* target = nullptr; int bkey_col = -1; std::vector<int> bpay_cols;
  // aggregate sink
  std::vector<int> group_cols; std::vector<PipeAgg> aggs; int agg_mode = DFGPU_AGG_SINGLE, agg_stage = -1, rows_word = -1; bool acc_ready = false;
  // output sink
  std::vector<int> out_cols; bool out_ordered = true;
  std::vector<std::vector<DCol>> out_parts; int64_t out_rows_pending = 0;
  int64_t batch_size = 0;
  bool finished = false;
  DevBuf params_dev, counters;
  std::deque<BatchPtr> outq;
  int64_t m_input_rows = 0, m_sink_rows = 0, m_output_rows = 0, m_groups = 0;
  std::string name;   // optional label: the kernel-timing family becomes "pipe:<name>" (dfgpu_kernel_time)
};

namespace dfgpu {

static LookupDev lookup_dev(const dfgpu_lookup* l) {
  LookupDev d;
  memset(&d, 0, sizeof(d));
  d.mode = l->mode; d.stride = l->stride; d.has_payload = l->has_payload ? 1 : 0;
  d.recs = l->recs.as<unsigned long long>(); d.cap = l->cap;
  d.bloom = l->bloom.ptr ? l->bloom.as<unsigned long long>() : nullptr; d.bloom_blocks = l->bloom_blocks;
  d.coarse = (l->bloom.ptr && l->coarse_words) ? (uint32_t*)(l->bloom.as<unsigned long long>() + l->bloom_blocks) : nullptr; d.coarse_words = l->coarse_words;
  d.bits = l->bits.ptr ? l->bits.as<uint32_t>() : nullptr; d.kmin = l->kmin; d.ksize = l->ksize;
  return d;
}

static bool key_type_ok(int t) { int w = type_width(t); return w >= 1 && w <= 8 && !type_is_float(t) && t != DFGPU_BOOL; }

// (re)allocate a hash lookup for at least `rows` records at load factor <= 0.5; existing records are rehashed
static void lookup_reserve(dfgpu_lookup* l, int64_t rows) {
  if (l->mode != LK_HASH || l->filter_only) return;
  dfgpu_ctx* ctx = l->ctx;
  const uint64_t need = std::max<uint64_t>(1024, (uint64_t)rows * 2);
  if (l->cap >= need) return;
  const uint64_t new_cap = l->cap == 0 ? need : std::max<uint64_t>(need, l->cap * 2);
  DF_CHECK(new_cap < 0xFFFFFFFFull, DFGPU_ERR_UNSUPPORTED, "lookup: more than 2^31 build rows");
  DevBuf nrecs(ctx, (size_t)new_cap * l->stride * 8), nbloom;
  lookup_init_kernel<<<grid_for((int64_t)new_cap * l->stride, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(nrecs.as<unsigned long long>(), new_cap, l->stride);
  DF_LAUNCH_CHECK(ctx);
  uint64_t nblocks = 0;
  const size_t table_bytes = (size_t)new_cap * l->stride * 8;
  if (l->opt.membership_filter == 1 || (l->opt.membership_filter < 0 && table_bytes > (96ull << 20))) {
    nblocks = std::max<uint64_t>(1024, new_cap / 8);   // 16 bits per key at load factor 0.5
    nbloom.alloc(ctx, (size_t)nblocks * 8);
    nbloom.zero();
  }
  LookupDev old_t = lookup_dev(l);
  LookupDev new_t = old_t;
  new_t.recs = nrecs.as<unsigned long long>(); new_t.cap = new_cap; new_t.bloom = nbloom.ptr ? nbloom.as<unsigned long long>() : nullptr; new_t.bloom_blocks = nblocks;
  if (l->cap > 0 && l->rows > 0) {
    lookup_rehash_kernel<<<grid_for((int64_t)l->cap, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(old_t, new_t);
    DF_LAUNCH_CHECK(ctx);
    l->rehashes++;
  }
  l->recs = std::move(nrecs); l->bloom = std::move(nbloom); l->cap = new_cap; l->bloom_blocks = nblocks;
}

static ColRef col_ref(const DCol& c) {
  ColRef r;
  r.ptr = c.values; r.valid = c.validity; r.voff = c.offset; r.width = type_width(c.type); r.sgn = type_is_signed_int(c.type) ? 1 : 0;
  r.vec = ((uintptr_t)c.values % 32 == 0) ? 2 : (((uintptr_t)c.values % 16 == 0) ? 1 : 0); r.pad = 0;   // 2: 256-bit loads allowed too
  return r;
}

// bind a planned expression into pool nodes; virtual columns >= n_cols become payload-field nodes
static int bind_pool(const dfgpu_pipeline* p, const ExprPlan& plan, const std::vector<DCol>& cols, PipeParams* pp, int* pool_used) {
  const int start = *pool_used;
  const int64_t n_rows = cols.empty() ? 0 : cols[0].length;
  const auto gmasks = resolve_guards(p->ctx, plan, cols, n_rows);   // short-circuit AND / OR: which RHS errors count on which rows (binary.rs:1182)
  DF_CHECK(start + (int)plan.nodes.size() <= kPoolNodes, DFGPU_ERR_UNSUPPORTED, "pipeline: expressions too large (56 nodes in total)");
  for (size_t i = 0; i < plan.nodes.size(); ++i) {
    const dfgpu_expr_node& nd = plan.nodes[i];
    ENode& e = pp->pool[start + i];
    memset(&e, 0, sizeof(e));
    e.kind = nd.kind; e.op = nd.a; e.in_type = plan.in_type[i]; e.out_type = plan.out_type[i];
    if (nd.kind == DFGPU_EXPR_COLUMN) {
      if (nd.a < (int)cols.size()) {
        const DCol& c = cols[nd.a];
        e.col = c.values; e.valid = c.validity; e.voff = c.offset;
      } else {
        const ExtDef& x = p->exts[nd.a - (int)cols.size()];
        e.kind = kExprExt; e.voff = x.stage; e.lit = (uint64_t)x.shift;
      }
    } else if (nd.kind == DFGPU_EXPR_LITERAL) {
      e.lit = literal_bits(nd); e.lit_null = nd.is_null;
      if (type_is_decimal(nd.type)) memcpy(&e.voff, &nd.lit_f64, 8);
    } else if ((nd.kind == DFGPU_EXPR_BINARY || nd.kind == DFGPU_EXPR_CAST) && plan.has_decimal) {
      e.voff = plan.aux[i];   // power-of-ten rescale exponents (expr_dec.cuh)
    }
    e.g_and = gmasks[i].first; e.g_or = gmasks[i].second;
  }
  *pool_used = start + (int)plan.nodes.size();
  return start;
}

// largest evaluation-stack depth of a post-order program (the register-resident interpreter handles <= 4)
static int plan_depth(const ExprPlan& plan) {
  int sp = 0, mx = 0;
  for (const auto& nd : plan.nodes) {
    if (nd.kind == DFGPU_EXPR_COLUMN || nd.kind == DFGPU_EXPR_LITERAL) sp++;
    else if (nd.kind == DFGPU_EXPR_BINARY) sp--;
    mx = std::max(mx, sp);
  }
  return mx;
}

// only integer columns without NULLs (in THIS batch), non-NULL integer literals, payload fields and + - *: eval_int_fast applies
static bool plan_is_int_arith(const dfgpu_pipeline* p, const ExprPlan& plan, const std::vector<DCol>& cols) {
  for (size_t i = 0; i < plan.nodes.size(); ++i) {
    const dfgpu_expr_node& nd = plan.nodes[i];
    const int t = plan.out_type[i];
    if (!type_is_int(t)) return false;
    if (nd.kind == DFGPU_EXPR_COLUMN) { if (nd.a < (int)cols.size() && cols[nd.a].validity) return false; }
    else if (nd.kind == DFGPU_EXPR_LITERAL) { if (nd.is_null) return false; }
    else if (nd.kind == DFGPU_EXPR_BINARY) { if (nd.a != DFGPU_OP_PLUS && nd.a != DFGPU_OP_MINUS && nd.a != DFGPU_OP_MULTIPLY) return false; }
    else return false;
  }
  (void)p;
  return true;
}

static bool expr_can_be_null(const ExprPlan& plan, const std::vector<DCol>& cols) {
  for (const auto& nd : plan.nodes) {
    if (nd.kind == DFGPU_EXPR_COLUMN && nd.a < (int)cols.size() && cols[nd.a].validity) return true;
    if (nd.kind == DFGPU_EXPR_LITERAL && nd.is_null) return true;
  }
  return false;
}

static void fill_params(dfgpu_pipeline* p, const std::vector<DCol>& cols, PipeParams* pp) {
  memset(pp, 0, sizeof(*pp));
  pp->n_cols = (int)cols.size();
  static const int hints_env = getenv("DFGPU_PIPE_HINTS") ? atoi(getenv("DFGPU_PIPE_HINTS")) : 3;
  pp->hints = hints_env;
  for (size_t c = 0; c < cols.size(); ++c) pp->col[c] = col_ref(cols[c]);
  int pool_used = 0;
  pp->pred_mode = 0;
  if (p->has_pred) {
    // fast path: conjunction of `column <cmp> literal` terms over integer-class columns
    const auto& nd = p->pred.nodes;
    std::vector<std::array<long long, 4>> terms;  // col, op, uns, lit
    bool fast = true;
    size_t i = 0;
    int depth = 0;
    while (i < nd.size() && fast) {
      if (i + 2 < nd.size() && nd[i].kind == DFGPU_EXPR_COLUMN && nd[i + 1].kind == DFGPU_EXPR_LITERAL && nd[i + 2].kind == DFGPU_EXPR_BINARY &&
          nd[i + 2].a >= DFGPU_OP_EQ && nd[i + 2].a <= DFGPU_OP_GTEQ && !nd[i + 1].is_null) {
        const int ct = p->in_types[nd[i].a];
        const int cls = cls_of(ct);
        if ((cls != C_I64 && cls != C_U64) || p->pred.has_decimal || (int)terms.size() >= kMaxTerms) { fast = false; break; }
        terms.push_back({(long long)nd[i].a, (long long)nd[i + 2].a, (long long)(cls == C_U64), (long long)literal_bits(nd[i + 1])});
        depth++; i += 3;
      } else if (nd[i].kind == DFGPU_EXPR_BINARY && nd[i].a == DFGPU_OP_AND && depth >= 2) { depth--; i++; }
      else fast = false;
    }
    if (fast && depth == 1 && !terms.empty()) {
      pp->pred_mode = 1; pp->n_terms = (int)terms.size();
      for (size_t t = 0; t < terms.size(); ++t) { pp->term_col[t] = (int)terms[t][0]; pp->term_op[t] = (int)terms[t][1]; pp->term_uns[t] = (int)terms[t][2]; pp->term_lit[t] = terms[t][3]; }
    } else {
      pp->pred_mode = 2;
      pp->pred_start = bind_pool(p, p->pred, cols, pp, &pool_used);
      pp->pred_n = (int)p->pred.nodes.size();
      pp->pred_small = p->pred.has_decimal ? 3 : (plan_depth(p->pred) <= 4 ? 1 : 0);
    }
  }
  pp->n_stages = (int)p->stages.size();
  pp->first_hash = -1;
  for (size_t s = 0; s < p->stages.size(); ++s) if (p->stages[s].lookup->mode == LK_HASH && p->stages[s].kind != DFGPU_STAGE_MAYBE && pp->first_hash < 0) pp->first_hash = (int)s;
  for (size_t s = 0; s < p->stages.size(); ++s) {
    pp->stage[s].kind = p->stages[s].kind; pp->stage[s].key_col = p->stages[s].key_col; pp->stage[s].lk = lookup_dev(p->stages[s].lookup);
  }
  pp->n_ext = (int)p->exts.size();
  for (size_t e = 0; e < p->exts.size(); ++e) pp->ext[e] = p->exts[e];
  pp->agg_stage = -1;
  if (p->sink == SINK_BUILD) {
    pp->target = lookup_dev(p->target);
    pp->bkey_col = p->bkey_col;
    pp->target_unique = (p->target->has_payload || p->target->opt.n_acc_words > 0) ? 1 : 0;
    pp->n_bpay = (int)p->bpay_cols.size();
    for (size_t c = 0; c < p->bpay_cols.size(); ++c) {
      pp->bpay_src[c] = p->bpay_cols[c]; pp->bpay_shift[c] = p->target->pay_shift[c]; pp->bpay_width[c] = type_width(p->target->pay_types[c]);
      if (p->bpay_cols[c] < (int)cols.size()) DF_CHECK(!cols[p->bpay_cols[c]].validity, DFGPU_ERR_UNSUPPORTED, "pipeline: nullable build payload columns stay on the unfused join");
    }
  } else if (p->sink == SINK_AGG) {
    pp->agg_stage = p->agg_stage; pp->rows_word = p->rows_word; pp->n_aggs = (int)p->aggs.size();
    for (size_t a = 0; a < p->aggs.size(); ++a) {
      const PipeAgg& ag = p->aggs[a];
      AggDef& d = pp->agg[a];
      d.func = ag.func; d.cls = ag.cls; d.word = ag.word; d.nn_word = ag.nn_word; d.start = 0; d.n = 0;
      if (ag.has_expr) {
        d.start = bind_pool(p, ag.plan, cols, pp, &pool_used); d.n = (int)ag.plan.nodes.size();
        d.small = plan_depth(ag.plan) <= 4 ? 1 : 0;
        if (d.small && plan_is_int_arith(p, ag.plan, cols)) d.small = 2;
        if (ag.plan.has_decimal) d.small = 3;
        if (ag.nn_word < 0 && ag.func != DFGPU_AGG_COUNT)
          DF_CHECK(!expr_can_be_null(ag.plan, cols), DFGPU_ERR_UNSUPPORTED, "pipeline: nullable aggregate input needs one more accumulator word in the lookup (n_acc_words)");
      }
    }
  }
}

static void check_errors(unsigned long long err) {
  if (err & ERR_DIV_ZERO) throw Error(DFGPU_ERR_ARITH, "Arrow error: Divide by zero error");
  if (err & ERR_OVERFLOW) throw Error(DFGPU_ERR_ARITH, "Arrow error: Arithmetic overflow");
  if (err & ERR_CAST) throw Error(DFGPU_ERR_ARITH, "Arrow error: Cast error: Can't cast value to the target type (out of range)");
}

template <int SINK>
static void launch_pipe(dfgpu_pipeline* p, int64_t n, const char* timer_name) {
  dfgpu_ctx* ctx = p->ctx;
  const int64_t ntiles = (n + kPipeTile - 1) / kPipeTile;
  // programs that touch Decimal128 values run a second instantiation of the kernel (128-bit interpreter linked in): the integer
  // instantiation stays byte for byte what it was
  bool dec = p->has_pred && p->pred.has_decimal;
  for (const auto& ag : p->aggs) dec = dec || (ag.has_expr && ag.plan.has_decimal);
  static const int blocks_env = getenv("DFGPU_PIPE_BLOCKS_PER_SM") ? atoi(getenv("DFGPU_PIPE_BLOCKS_PER_SM")) : 0;
  int blocks_per_sm = blocks_env;
  if (blocks_per_sm <= 0) {   // persistent blocks: exactly one resident wave (a second wave would start after the first finished)
    if (dec) DF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, pipe_kernel<SINK, true>, kPipeThreads, 0));
    else DF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, pipe_kernel<SINK, false>, kPipeThreads, 0));
    blocks_per_sm = std::max(1, blocks_per_sm);
  }
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)kNumSMs * blocks_per_sm);
  const std::string tname = p->name.empty() ? std::string(timer_name) : "pipe:" + p->name;
  KernelTimer kt(ctx, tname.c_str());
  const int var_env = getenv("DFGPU_PIPE_VAR") ? atoi(getenv("DFGPU_PIPE_VAR")) : kPipeVarDefault;
  const PipeParams* gp = (const PipeParams*)p->params_dev.ptr;
  unsigned long long* cnt = p->counters.as<unsigned long long>();
  if (dec) pipe_kernel<SINK, true><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 1) pipe_kernel<SINK_AGG, false, 1><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 2) pipe_kernel<SINK_AGG, false, 2><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 3) pipe_kernel<SINK_AGG, false, 3><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 8) pipe_kernel<SINK_AGG, false, 8><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 9) pipe_kernel<SINK_AGG, false, 9><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 11) pipe_kernel<SINK_AGG, false, 11><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_AGG && var_env == 43) pipe_kernel<SINK_AGG, false, 43><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_PACK && (var_env & 32)) pipe_kernel<SINK_PACK, false, 34><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_PACK && (var_env & 2)) pipe_kernel<SINK_PACK, false, 2><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  else if (SINK == SINK_OUTPUT_ANY && (var_env & 2)) pipe_kernel<SINK_OUTPUT_ANY, false, 2><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);   // the multi-GPU plan's scans (the 256-bit loads were not measured on this sink)
  else pipe_kernel<SINK, false><<<grid, kPipeThreads, 0, ctx->stream>>>(gp, n, cnt);
  DF_LAUNCH_CHECK(ctx);
}

static void read_counters(dfgpu_pipeline* p, unsigned long long h[4]) {
  DF_CUDA(cudaMemcpyAsync(h, p->counters.ptr, 32, cudaMemcpyDeviceToHost, p->ctx->stream));
  DF_CUDA(cudaStreamSynchronize(p->ctx->stream));
}

static void upload_params(dfgpu_pipeline* p, const PipeParams& pp) {
  if (!p->params_dev.ptr) p->params_dev.alloc(p->ctx, sizeof(PipeParams));
  DF_CUDA(cudaMemcpyAsync(p->params_dev.ptr, &pp, sizeof(PipeParams), cudaMemcpyHostToDevice, p->ctx->stream));
  DF_CUDA(cudaStreamSynchronize(p->ctx->stream));   // `pp` lives on the caller's stack frame
}

static void prepare_acc(dfgpu_pipeline* p) {
  if (p->acc_ready) return;
  dfgpu_lookup* l = p->stages[p->agg_stage].lookup;
  dfgpu_ctx* ctx = p->ctx;
  LookupDev t = lookup_dev(l);
  for (const PipeAgg& ag : p->aggs) {
    if (ag.func != DFGPU_AGG_MIN && ag.func != DFGPU_AGG_MAX) continue;
    unsigned long long init;
    const bool is_min = ag.func == DFGPU_AGG_MIN;
    if (ag.cls == C_F64) { double d = is_min ? INFINITY : -INFINITY; memcpy(&init, &d, 8); }
    else if (ag.cls == C_U64) init = is_min ? ~0ull : 0ull;
    else init = is_min ? (unsigned long long)LLONG_MAX : (unsigned long long)LLONG_MIN;
    if (l->cap) { lookup_init_acc_kernel<<<grid_for((int64_t)l->cap, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(t, ag.word, init); DF_LAUNCH_CHECK(ctx); }
  }
  p->acc_ready = true;
}

static void pipeline_push(dfgpu_pipeline* p, const std::vector<DCol>& cols) {
  DF_CHECK(!p->finished, DFGPU_ERR_STATE, "push after finish");
  DF_CHECK(p->sink != SINK_NONE, DFGPU_ERR_STATE, "pipeline: choose a sink before the first push");
  DF_CHECK(cols.size() == p->in_types.size(), DFGPU_ERR_INVALID, "pipeline input column count mismatch");
  dfgpu_ctx* ctx = p->ctx;
  set_device(ctx);
  const int64_t n = cols.empty() ? 0 : cols[0].length;
  for (size_t c = 0; c < cols.size(); ++c) {
    DF_CHECK(cols[c].type == p->in_types[c], DFGPU_ERR_INVALID, "pipeline input column type mismatch");
    DF_CHECK(cols[c].length == n, DFGPU_ERR_INVALID, "pipeline input ragged columns");
    DF_CHECK(cols[c].type != DFGPU_BOOL && (type_width(cols[c].type) <= 8 || type_is_decimal(cols[c].type)), DFGPU_ERR_UNSUPPORTED,
             "pipeline: fixed-width columns of <= 8 bytes (and Decimal128 inside expressions) only");
  }
  p->m_input_rows += n;
  if (n == 0) return;
  if (!p->counters.ptr) p->counters.alloc(ctx, 64);
  PipeParams pp;
  unsigned long long h[4];
  if (p->sink == SINK_BUILD) {
    dfgpu_lookup* t = p->target;
    if (t->mode == LK_HASH && !t->filter_only && (uint64_t)(t->rows + n) * 2 > t->cap) {
      // the batch may not fit at load factor 0.5 and nobody knows how many rows survive: ONE pass evaluates the pipeline and leaves
      // the survivors as packed {key, payload} records; the table is sized for exactly that many and the records are inserted by a
      // dense kernel (no second scan of the input).
      fill_params(p, cols, &pp);
      DevBuf recs(ctx, (size_t)n * 16);
      p->counters.zero();
      pp.out_dst[0] = recs.ptr;
      pp.out_counter = p->counters.as<unsigned long long>() + 4;
      upload_params(p, pp);
      launch_pipe<SINK_PACK>(p, n, "pipeline_build");
      unsigned long long h8[8];
      DF_CUDA(cudaMemcpyAsync(h8, p->counters.ptr, 64, cudaMemcpyDeviceToHost, ctx->stream));
      DF_CUDA(cudaStreamSynchronize(ctx->stream));
      check_errors(h8[3]);
      const int64_t packed = (int64_t)h8[4];
      lookup_reserve(t, t->rows + packed);
      p->counters.zero();
      if (packed > 0) {
        KernelTimer kt(ctx, "lookup_insert");
        lookup_insert_records_kernel<<<grid_for(packed, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(lookup_dev(t), (const ulonglong2*)recs.ptr, packed,
                                                                                                   (t->has_payload || t->opt.n_acc_words > 0) ? 1 : 0, p->counters.as<unsigned long long>());
        DF_LAUNCH_CHECK(ctx);
      }
      read_counters(p, h);
      h[0] = h8[0];
    } else {
      fill_params(p, cols, &pp);
      upload_params(p, pp);
      p->counters.zero();
      launch_pipe<SINK_BUILD>(p, n, "pipeline_build");
      read_counters(p, h);
      check_errors(h[3]);
    }
    if (h[2] & 2) throw Error(DFGPU_ERR_INVALID, "lookup build: a key lies outside the declared key range (or is the reserved all-ones value)");
    if (h[2] & 1) throw Error(DFGPU_ERR_UNSUPPORTED, "lookup build: duplicate build keys — the fused lookup needs unique keys, use dfgpu_hashjoin");
    t->rows += (int64_t)h[1];
    p->m_sink_rows += (int64_t)h[0];
  } else if (p->sink == SINK_AGG) {
    prepare_acc(p);
    fill_params(p, cols, &pp);
    upload_params(p, pp);
    p->counters.zero();
    launch_pipe<SINK_AGG>(p, n, "pipeline_agg");
    read_counters(p, h);
    check_errors(h[3]);
    p->m_sink_rows += (int64_t)h[0];
  } else if (!p->out_ordered) {   // SINK_OUTPUT, row order unspecified: the two-phase kernel, one global reservation per 128 survivors
    DF_CHECK(n < 0xFFFFFFFFll, DFGPU_ERR_UNSUPPORTED, "pipeline: a batch must have < 2^32-1 rows");
    fill_params(p, cols, &pp);
    std::vector<DCol> part;
    pp.n_out = (int)p->out_cols.size();
    for (int c = 0; c < pp.n_out; ++c) {
      const int src = p->out_cols[c];
      if (src < (int)cols.size()) DF_CHECK(!cols[src].validity, DFGPU_ERR_UNSUPPORTED, "pipeline output: nullable columns stay on the unfused operators");
      DCol d = alloc_col(ctx, p->vtypes[src], n, false);
      pp.out_src[c] = src; pp.out_width[c] = type_width(p->vtypes[src]); pp.out_dst[c] = d.own_values->ptr;
      part.push_back(std::move(d));
    }
    p->counters.zero();
    pp.out_counter = p->counters.as<unsigned long long>() + 4;
    upload_params(p, pp);
    launch_pipe<SINK_OUTPUT_ANY>(p, n, "pipeline_output");
    unsigned long long h8[8];
    DF_CUDA(cudaMemcpyAsync(h8, p->counters.ptr, 64, cudaMemcpyDeviceToHost, ctx->stream));
    DF_CUDA(cudaStreamSynchronize(ctx->stream));
    check_errors(h8[3]);
    const int64_t kept = (int64_t)h8[4];
    p->m_sink_rows += kept;
    if (kept > 0) {
      for (auto& c : part) c.length = kept;
      p->out_parts.push_back(std::move(part));
      p->out_rows_pending += kept;
    }
  } else {  // SINK_OUTPUT, input order preserved
    DF_CHECK(n < 0xFFFFFFFFll, DFGPU_ERR_UNSUPPORTED, "pipeline: a batch must have < 2^32-1 rows");
    for (auto& st : p->stages) DF_CHECK(st.kind != DFGPU_STAGE_MAYBE, DFGPU_ERR_UNSUPPORTED, "pipeline: MAYBE stages feed an exchange — use the unordered output sink");
    fill_params(p, cols, &pp);
    upload_params(p, pp);
    p->counters.zero();
    OutCols oc;
    memset(&oc, 0, sizeof(oc));
    oc.n = (int)p->out_cols.size();
    std::vector<DCol> part;
    for (int c = 0; c < oc.n; ++c) {
      const int src = p->out_cols[c];
      if (src < (int)cols.size()) DF_CHECK(!cols[src].validity, DFGPU_ERR_UNSUPPORTED, "pipeline output: nullable columns stay on the unfused operators");
      DCol d = alloc_col(ctx, p->vtypes[src], n, false);
      oc.src[c] = src; oc.width[c] = type_width(p->vtypes[src]); oc.dst[c] = d.own_values->ptr;
      part.push_back(std::move(d));
    }
    const int64_t nt = (n + kPipeTile - 1) / kPipeTile;
    DevBuf desc(ctx, (size_t)nt * 8 + 32);
    desc.zero();
    unsigned long long* totals = (unsigned long long*)((char*)desc.ptr + (size_t)nt * 8);
    unsigned int* counter = (unsigned int*)(totals + 2);
    {
      KernelTimer kt(ctx, "pipeline_output");
      pipe_output_kernel<<<(int)nt, kPipeThreads, 0, ctx->stream>>>((const PipeParams*)p->params_dev.ptr, n, oc, desc.as<unsigned long long>(), counter, totals, p->counters.as<unsigned long long>());
      DF_LAUNCH_CHECK(ctx);
    }
    unsigned long long tot[2];
    DF_CUDA(cudaMemcpyAsync(tot, totals, 16, cudaMemcpyDeviceToHost, ctx->stream));
    read_counters(p, h);
    check_errors(h[3]);
    const int64_t kept = (int64_t)tot[0];
    p->m_sink_rows += kept;
    if (kept > 0) {
      for (auto& c : part) c.length = kept;
      p->out_parts.push_back(std::move(part));
      p->out_rows_pending += kept;
    }
  }
}

static void emit_sliced(dfgpu_pipeline* p, std::vector<DCol>& merged, int64_t rows) {
  const int64_t bs = p->batch_size > 0 ? p->batch_size : std::max<int64_t>(rows, 1);
  for (int64_t pos = 0; pos < rows; pos += bs) {
    const int64_t len = std::min<int64_t>(bs, rows - pos);
    BatchPtr b(new dfgpu_batch());
    b->ctx = p->ctx; b->rows = len; b->host = false;
    for (auto& c : merged) b->cols.push_back((pos == 0 && len == rows) ? c : slice_column(c, pos, len));
    p->m_output_rows += len;
    p->outq.push_back(std::move(b));
  }
}

static void pipeline_finish(dfgpu_pipeline* p) {
  DF_CHECK(!p->finished, DFGPU_ERR_STATE, "finish called twice");
  p->finished = true;
  dfgpu_ctx* ctx = p->ctx;
  set_device(ctx);
  if (p->sink == SINK_OUTPUT) {
    if (p->out_rows_pending == 0) return;
    std::vector<DCol> merged;
    for (size_t c = 0; c < p->out_cols.size(); ++c) {
      std::vector<DCol> parts;
      for (auto& b : p->out_parts) parts.push_back(b[c]);
      merged.push_back(parts.size() == 1 ? parts[0] : concat_columns(ctx, parts, p->vtypes[p->out_cols[c]]));
    }
    p->out_parts.clear();
    emit_sliced(p, merged, p->out_rows_pending);
    return;
  }
  if (p->sink != SINK_AGG) return;
  dfgpu_lookup* l = p->stages[p->agg_stage].lookup;
  if (l->cap == 0) return;
  LookupDev t = lookup_dev(l);
  const uint64_t nw = (l->cap + 31) / 32;
  DevBuf words(ctx, (size_t)nw * 4 + 8), idx;
  lookup_groups_kernel<<<grid_for((int64_t)nw * 32, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(t, p->rows_word, words.as<uint32_t>());
  DF_LAUNCH_CHECK(ctx);
  const int64_t groups = compact_flag_indices(ctx, words.as<uint32_t>(), (int64_t)l->cap, 1, &idx);
  p->m_groups = groups;
  if (groups == 0) return;
  // output schema: group columns, then one (Single) or the state (Partial) columns per aggregate
  const bool partial = p->agg_mode == DFGPU_AGG_PARTIAL;
  EmitCols ec;
  memset(&ec, 0, sizeof(ec));
  std::vector<DCol> out;
  const int key_col = p->stages[p->agg_stage].key_col;
  auto add = [&](int type, bool nullable, EmitCol e) {
    DF_CHECK(ec.n < kMaxPipeCols, DFGPU_ERR_UNSUPPORTED, "pipeline: too many output columns");
    DCol d = alloc_col(ctx, type, groups, nullable);
    e.width = type_width(type); e.dst = d.own_values->ptr; e.valid = nullable ? d.own_validity->as<uint32_t>() : nullptr;
    d.null_count = nullable ? -1 : 0;
    ec.c[ec.n++] = e;
    out.push_back(std::move(d));
  };
  for (int g : p->group_cols) {
    EmitCol e; memset(&e, 0, sizeof(e));
    if (g == key_col) { e.kind = 0; add(p->in_types[g], false, e); }
    else { const ExtDef& x = p->exts[g - (int)p->in_types.size()]; e.kind = 1; e.shift = x.shift; add(x.type, false, e); }
  }
  for (const PipeAgg& ag : p->aggs) {
    EmitCol e; memset(&e, 0, sizeof(e));
    e.nn_word = -1;
    const int sum_type = ag.cls == C_F64 ? DFGPU_FLOAT64 : (ag.cls == C_U64 ? DFGPU_UINT64 : DFGPU_INT64);   // sum.rs:232-261
    switch (ag.func) {
      case DFGPU_AGG_COUNT_STAR: e.kind = 4; e.word = p->rows_word; add(DFGPU_INT64, false, e); break;
      case DFGPU_AGG_COUNT: e.kind = 4; e.word = ag.word; add(DFGPU_INT64, false, e); break;
      case DFGPU_AGG_SUM:
        if (ag.cls == C_DEC) {   // Sum::return_type: Decimal128(min(38, p + 10), s) (sum.rs:247-249); the Partial state has the same type
          e.kind = 5; e.word = ag.word; e.nn_word = ag.nn_word;
          add(dec_type(std::min(38, dec_precision(ag.arg_type) + 10), dec_scale(ag.arg_type)), ag.nn_word >= 0, e);
        } else { e.kind = 2; e.word = ag.word; e.nn_word = ag.nn_word; add(sum_type, ag.nn_word >= 0, e); }
        break;
      case DFGPU_AGG_MIN: case DFGPU_AGG_MAX: e.kind = 2; e.word = ag.word; e.nn_word = ag.nn_word; add(ag.arg_type, ag.nn_word >= 0, e); break;
      case DFGPU_AGG_AVG:
        if (partial) {   // state = [count: UInt64, sum: Float64] (aggregates/mod.rs:3591-3700)
          EmitCol c1 = e; c1.kind = 4; c1.word = ag.cnt_word; add(DFGPU_UINT64, false, c1);
          EmitCol c2 = e; c2.kind = 2; c2.word = ag.word; c2.nn_word = ag.cnt_word; add(DFGPU_FLOAT64, true, c2);
        } else { e.kind = 3; e.word = ag.word; e.cnt_word = ag.cnt_word; add(DFGPU_FLOAT64, true, e); }
        break;
    }
  }
  lookup_emit_kernel<<<grid_for(groups, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(t, idx.as<uint32_t>(), groups, p->rows_word, ec);
  DF_LAUNCH_CHECK(ctx);
  emit_sliced(p, out, groups);
}

}  // namespace dfgpu

extern "C" {

void dfgpu_lookup_default_options(dfgpu_lookup_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->membership_filter = -1;
}

int dfgpu_lookup_create(dfgpu_ctx* ctx, int32_t key_type, const int32_t* payload_types, int32_t n_payload, const dfgpu_lookup_options* opts, dfgpu_lookup** out) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && out, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(key_type_ok(key_type), DFGPU_ERR_UNSUPPORTED, "lookup: the key must be one integer-like column of <= 64 bits");
  DF_CHECK(n_payload >= 0 && n_payload <= kMaxBuildPay && (n_payload == 0 || payload_types), DFGPU_ERR_INVALID, "lookup: 0..8 payload columns");
  set_device(ctx);
  std::unique_ptr<dfgpu_lookup> l(new dfgpu_lookup());
  l->ctx = ctx; l->key_type = key_type;
  if (opts) l->opt = *opts; else dfgpu_lookup_default_options(&l->opt);
  DF_CHECK(l->opt.n_acc_words >= 0 && l->opt.n_acc_words <= 12, DFGPU_ERR_INVALID, "lookup: 0..12 accumulator words");
  int bits = 0;
  for (int i = 0; i < n_payload; ++i) {
    const int w = type_width(payload_types[i]);
    DF_CHECK(w >= 1 && w <= 8 && payload_types[i] != DFGPU_BOOL, DFGPU_ERR_UNSUPPORTED, "lookup: payload columns must be fixed-width, <= 8 bytes");
    l->pay_types.push_back(payload_types[i]); l->pay_shift.push_back(bits);
    bits += 8 * w;
  }
  DF_CHECK(bits <= 64, DFGPU_ERR_UNSUPPORTED, "lookup: payload columns wider than 64 bits together stay on the unfused join");
  l->has_payload = n_payload > 0;
  l->stride = 1 + (l->has_payload ? 1 : 0) + l->opt.n_acc_words;
  if (l->has_payload && (l->stride & 1)) l->stride++;   // {key,payload} is claimed with one 16-byte CAS
  l->mode = LK_HASH;
  if (!l->has_payload && l->opt.n_acc_words == 0 && l->opt.has_key_range && l->opt.key_max >= l->opt.key_min) {
    const uint64_t range = (uint64_t)l->opt.key_max - (uint64_t)l->opt.key_min;
    if (range < (1ull << 32)) {   // <= 512 MiB of bits
      l->mode = LK_BITMAP; l->kmin = (uint64_t)l->opt.key_min; l->ksize = range + 1;
      l->bits.alloc(ctx, (size_t)((l->ksize + 31) / 32) * 4 + 8);
      l->bits.zero();
    }
  }
  if (l->opt.filter_only) {
    DF_CHECK(l->mode == LK_HASH && !l->has_payload && l->opt.n_acc_words == 0, DFGPU_ERR_INVALID, "lookup: a filter-only lookup is a key set without payload");
    DF_CHECK(l->opt.expected_rows > 0, DFGPU_ERR_INVALID, "lookup: a filter-only lookup needs expected_rows (its geometry is fixed up front)");
    l->filter_only = true;
    l->bloom_blocks = std::max<uint64_t>(1024, (uint64_t)l->opt.expected_rows / 4);   // 16 bits per key
    DF_CHECK(l->bloom_blocks < (1ull << 32), DFGPU_ERR_UNSUPPORTED, "lookup: filter too large");
    // a filter larger than L2 gets the coarse first level (4 bits per key) in the same allocation: one buffer to share and to all-reduce
    static const int coarse_mb = getenv("DFGPU_FILTER_COARSE_MB") ? atoi(getenv("DFGPU_FILTER_COARSE_MB")) : 48;
    if (coarse_mb >= 0 && (size_t)l->bloom_blocks * 8 > ((size_t)coarse_mb << 20)) l->coarse_words = (std::max<uint64_t>(4096, (uint64_t)l->opt.expected_rows / 8) + 1) & ~1ull;
    l->bloom.alloc(ctx, (size_t)l->bloom_blocks * 8 + (size_t)l->coarse_words * 4);
    l->bloom.zero();
  }
  if (l->mode == LK_HASH && !l->filter_only && l->opt.expected_rows > 0) lookup_reserve(l.get(), l->opt.expected_rows);
  *out = l.release();
  DF_API_END
}
int64_t dfgpu_lookup_metric(dfgpu_lookup* l, const char* name) {
  if (!l || !name) return -1;
  std::string s(name);
  if (s == "rows") return l->rows;
  if (s == "capacity") return l->mode == LK_BITMAP ? (int64_t)l->ksize : (int64_t)l->cap;
  if (s == "mode") return l->mode;
  if (s == "table_bytes") return l->mode == LK_BITMAP ? (int64_t)l->bits.bytes : (int64_t)l->recs.bytes;
  if (s == "filter_bytes") return (int64_t)l->bloom.bytes;
  if (s == "rehashes") return l->rehashes;
  if (s == "stride_bytes") return l->stride * 8;
  return -1;
}
int dfgpu_lookup_clear(dfgpu_lookup* l) {
  DF_API_BEGIN(l ? l->ctx : nullptr)
  DF_CHECK(l, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(!l->acc_claimed, DFGPU_ERR_STATE, "lookup: accumulators in use by a pipeline");
  dfgpu_ctx* ctx = l->ctx;
  set_device(ctx);
  if (l->mode == LK_BITMAP) l->bits.zero();
  else {
    if (l->cap) { lookup_init_kernel<<<grid_for((int64_t)l->cap * l->stride, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(l->recs.as<unsigned long long>(), l->cap, l->stride); DF_LAUNCH_CHECK(ctx); }
    if (l->bloom.ptr) l->bloom.zero();
  }
  l->rows = 0;
  DF_API_END
}
int dfgpu_lookup_filter_buffer(dfgpu_lookup* l, void** words_dev, uint64_t* n_bytes) {
  DF_API_BEGIN(l ? l->ctx : nullptr)
  DF_CHECK(l && words_dev && n_bytes, DFGPU_ERR_INVALID, "null argument");
  *words_dev = l->bloom.ptr; *n_bytes = (uint64_t)l->bloom_blocks * 8 + (uint64_t)l->coarse_words * 4;
  DF_API_END
}
int dfgpu_lookup_filter_allreduce_peer(dfgpu_lookup* l, void* const* peer_words, int32_t rank, int32_t n_ranks) {
  DF_API_BEGIN(l ? l->ctx : nullptr)
  DF_CHECK(l && peer_words, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks, DFGPU_ERR_INVALID, "filter all-reduce: 1..8 ranks of one box");
  DF_CHECK(l->bloom.ptr && l->bloom_blocks > 0, DFGPU_ERR_STATE, "filter all-reduce: the lookup has no membership filter");
  dfgpu_ctx* ctx = l->ctx;
  set_device(ctx);
  PeerWords pw;
  memset(&pw, 0, sizeof(pw));
  for (int q = 0; q < n_ranks; ++q) { DF_CHECK(peer_words[q], DFGPU_ERR_INVALID, "filter all-reduce: null peer pointer"); pw.p[q] = (unsigned long long*)peer_words[q]; }
  DF_CHECK(pw.p[rank] == l->bloom.as<unsigned long long>(), DFGPU_ERR_INVALID, "filter all-reduce: peer_words[rank] must be this lookup's own filter");
  const uint64_t words64 = l->bloom_blocks + l->coarse_words / 2;   // exact blocks + coarse level, merged as one array of 64-bit words
  const uint64_t slice = words64 / (uint64_t)n_ranks + 1;
  KernelTimer kt(ctx, "filter_allreduce");
  filter_allreduce_peer_kernel<<<grid_for((int64_t)slice, 256, kNumSMs * 4), 256, 0, ctx->stream>>>(pw, rank, n_ranks, words64);
  DF_LAUNCH_CHECK(ctx);
  DF_API_END
}
void dfgpu_lookup_destroy(dfgpu_lookup* l) {
  if (!l) return;
  cudaSetDevice(l->ctx->device);
  delete l;
}

int dfgpu_column_minmax_device(dfgpu_ctx* ctx, const dfgpu_column* col, int64_t* min_out, int64_t* max_out, int64_t* valid_out) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && col && min_out && max_out, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(key_type_ok(col->type), DFGPU_ERR_UNSUPPORTED, "minmax: integer-like columns only");
  set_device(ctx);
  DCol c = device_view(*col);
  const int uns = type_is_unsigned_int(c.type) ? 1 : 0;
  DevBuf mm(ctx, 24);
  unsigned long long init[3] = {~0ull, 0ull, 0ull};
  DF_CUDA(cudaMemcpyAsync(mm.ptr, init, 24, cudaMemcpyHostToDevice, ctx->stream));
  if (c.length > 0) {
    col_minmax_kernel<<<grid_for(c.length, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(col_ref(c), c.length, uns, mm.as<unsigned long long>());
    DF_LAUNCH_CHECK(ctx);
  }
  unsigned long long h[3];
  DF_CUDA(cudaMemcpyAsync(h, mm.ptr, 24, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  if (valid_out) *valid_out = (int64_t)h[2];
  if (h[2] == 0) { *min_out = 0; *max_out = -1; }
  else if (uns) { *min_out = (int64_t)h[0]; *max_out = (int64_t)h[1]; }
  else { *min_out = (int64_t)(h[0] ^ (1ull << 63)); *max_out = (int64_t)(h[1] ^ (1ull << 63)); }
  DF_API_END
}

int dfgpu_column_sum_device(dfgpu_ctx* ctx, const dfgpu_column* col, uint64_t* sum_out, int64_t* valid_out) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && col && sum_out, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(key_type_ok(col->type) || type_is_decimal(col->type), DFGPU_ERR_UNSUPPORTED, "column sum: integer-like and Decimal128 columns only");
  set_device(ctx);
  DCol c = device_view(*col);
  DevBuf acc(ctx, 16);
  acc.zero();
  if (c.length > 0) {
    col_sum_kernel<<<grid_for(c.length, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(col_ref(c), c.length, acc.as<unsigned long long>());
    DF_LAUNCH_CHECK(ctx);
  }
  unsigned long long h[2];
  DF_CUDA(cudaMemcpyAsync(h, acc.ptr, 16, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  *sum_out = h[0];
  if (valid_out) *valid_out = (int64_t)h[1];
  DF_API_END
}

int dfgpu_pipeline_create(dfgpu_ctx* ctx, const int32_t* input_types, int32_t n_cols, const dfgpu_expr_node* predicate, int32_t n_pred_nodes,
                          const dfgpu_pipeline_stage* stages, int32_t n_stages, dfgpu_pipeline** out) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && out && input_types, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(n_cols >= 1 && n_cols <= kMaxPipeCols, DFGPU_ERR_UNSUPPORTED, "pipeline: 1..16 input columns");
  DF_CHECK(n_stages >= 0 && n_stages <= kMaxStages && (n_stages == 0 || stages), DFGPU_ERR_UNSUPPORTED, "pipeline: 0..3 probe stages");
  std::unique_ptr<dfgpu_pipeline> p(new dfgpu_pipeline());
  p->ctx = ctx;
  p->in_types.assign(input_types, input_types + n_cols);
  p->vtypes = p->in_types;
  if (predicate && n_pred_nodes > 0) {
    p->pred = plan_expr(input_types, n_cols, predicate, n_pred_nodes);
    DF_CHECK(p->pred.root_type == DFGPU_BOOL, DFGPU_ERR_INVALID, "Cannot create filter_array from non-boolean predicates");
    p->has_pred = true;
  }
  for (int s = 0; s < n_stages; ++s) {
    const dfgpu_pipeline_stage& st = stages[s];
    DF_CHECK(st.lookup, DFGPU_ERR_INVALID, "pipeline: stage without a lookup");
    DF_CHECK(st.kind >= DFGPU_STAGE_INNER && st.kind <= DFGPU_STAGE_MAYBE, DFGPU_ERR_INVALID, "pipeline: unknown stage kind");
    DF_CHECK(!st.lookup->filter_only || st.kind == DFGPU_STAGE_MAYBE, DFGPU_ERR_INVALID, "pipeline: a filter-only lookup can only back a MAYBE stage");
    DF_CHECK(st.key_col >= 0 && st.key_col < n_cols, DFGPU_ERR_INVALID, "pipeline: stage key column out of range");
    const int kt = input_types[st.key_col], lt = st.lookup->key_type;
    DF_CHECK(key_type_ok(kt) && type_width(kt) == type_width(lt) && type_is_signed_int(kt) == type_is_signed_int(lt), DFGPU_ERR_INVALID,
             "pipeline: probe key type differs from the lookup's key type");
    DF_CHECK(st.lookup->ctx->device == ctx->device, DFGPU_ERR_INVALID, "pipeline: lookup lives on another device");
    p->stages.push_back(st);
    if (st.kind == DFGPU_STAGE_INNER)
      for (size_t f = 0; f < st.lookup->pay_types.size(); ++f) {
        DF_CHECK(p->exts.size() < (size_t)kMaxExt, DFGPU_ERR_UNSUPPORTED, "pipeline: at most 8 payload fields");
        ExtDef e; e.stage = s; e.shift = st.lookup->pay_shift[f]; e.width = type_width(st.lookup->pay_types[f]); e.type = st.lookup->pay_types[f];
        p->exts.push_back(e);
        p->vtypes.push_back(e.type);
      }
  }
  *out = p.release();
  DF_API_END
}

int dfgpu_pipeline_sink_build(dfgpu_pipeline* p, dfgpu_lookup* target, int32_t key_col, const int32_t* payload_cols, int32_t n_payload) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p && target, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(p->sink == SINK_NONE && p->m_input_rows == 0, DFGPU_ERR_STATE, "pipeline: the sink is chosen once, before the first push");
  DF_CHECK(key_col >= 0 && key_col < (int)p->in_types.size(), DFGPU_ERR_INVALID, "pipeline build sink: the key must be an input column");
  const int kt = p->in_types[key_col];
  DF_CHECK(type_width(kt) == type_width(target->key_type) && type_is_signed_int(kt) == type_is_signed_int(target->key_type), DFGPU_ERR_INVALID,
           "pipeline build sink: key type differs from the lookup's key type");
  DF_CHECK(n_payload == (int)target->pay_types.size(), DFGPU_ERR_INVALID, "pipeline build sink: payload column count differs from the lookup's");
  for (int c = 0; c < n_payload; ++c) {
    DF_CHECK(payload_cols[c] >= 0 && payload_cols[c] < (int)p->vtypes.size(), DFGPU_ERR_INVALID, "pipeline build sink: payload column out of range");
    DF_CHECK(type_width(p->vtypes[payload_cols[c]]) == type_width(target->pay_types[c]), DFGPU_ERR_INVALID, "pipeline build sink: payload column width differs from the lookup's");
    p->bpay_cols.push_back(payload_cols[c]);
  }
  for (auto& st : p->stages) DF_CHECK(st.lookup != target, DFGPU_ERR_INVALID, "pipeline: cannot build the lookup it probes");
  p->target = target; p->bkey_col = key_col; p->sink = SINK_BUILD;
  DF_API_END
}

int dfgpu_pipeline_sink_aggregate(dfgpu_pipeline* p, const int32_t* group_cols, int32_t n_group, const dfgpu_pipeline_agg* aggs, int32_t n_aggs, int32_t mode, int64_t batch_size) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p && group_cols && n_group >= 1, DFGPU_ERR_INVALID, "null argument");
  DF_CHECK(p->sink == SINK_NONE && p->m_input_rows == 0, DFGPU_ERR_STATE, "pipeline: the sink is chosen once, before the first push");
  DF_CHECK(n_aggs >= 0 && n_aggs <= kMaxPipeAggs && (n_aggs == 0 || aggs), DFGPU_ERR_UNSUPPORTED, "pipeline: 0..4 aggregates");
  DF_CHECK(mode == DFGPU_AGG_SINGLE || mode == DFGPU_AGG_SINGLE_PARTITIONED || mode == DFGPU_AGG_PARTIAL, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: Single / SinglePartitioned / Partial");
  // functional dependence: every group column is the probe key of ONE inner stage or a payload field of that stage
  const int nin = (int)p->in_types.size();
  int stage = -1;
  for (size_t s = 0; s < p->stages.size() && stage < 0; ++s) {
    if (p->stages[s].kind != DFGPU_STAGE_INNER) continue;
    bool has_key = false, ok = true;
    for (int g = 0; g < n_group; ++g) {
      const int c = group_cols[g];
      DF_CHECK(c >= 0 && c < (int)p->vtypes.size(), DFGPU_ERR_INVALID, "pipeline aggregate: group column out of range");
      if (c == p->stages[s].key_col) has_key = true;
      else if (c >= nin && p->exts[c - nin].stage == (int)s) {}
      else ok = false;
    }
    if (ok && has_key) stage = (int)s;
  }
  DF_CHECK(stage >= 0, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: group keys are not determined by one join key — use the unfused dfgpu_agg");
  dfgpu_lookup* l = p->stages[stage].lookup;
  DF_CHECK(!l->acc_claimed, DFGPU_ERR_STATE, "pipeline aggregate: the lookup's accumulators are already in use");
  const int base = 1 + (l->has_payload ? 1 : 0);
  int next = base, budget = base + l->opt.n_acc_words;
  DF_CHECK(next < budget, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: the lookup reserves no accumulator words (n_acc_words)");
  const int rows_word = next++;
  std::vector<PipeAgg> new_aggs;   // committed only when every check has passed
  for (int a = 0; a < n_aggs; ++a) {
    PipeAgg ag;
    ag.func = aggs[a].func;
    DF_CHECK(ag.func >= DFGPU_AGG_SUM && ag.func <= DFGPU_AGG_COUNT_STAR, DFGPU_ERR_INVALID, "pipeline aggregate: unknown function");
    if (ag.func != DFGPU_AGG_COUNT_STAR) {
      DF_CHECK(aggs[a].expr && aggs[a].n_nodes > 0, DFGPU_ERR_INVALID, "pipeline aggregate: missing argument expression");
      ag.plan = plan_expr(p->vtypes.data(), (int)p->vtypes.size(), aggs[a].expr, aggs[a].n_nodes);
      ag.has_expr = true;
      ag.arg_type = ag.plan.root_type;
      ag.cls = cls_of(ag.arg_type);
      DF_CHECK(ag.cls != C_BOOL || ag.func == DFGPU_AGG_COUNT, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: Boolean arguments only for COUNT");
      if (ag.func == DFGPU_AGG_AVG) DF_CHECK(ag.arg_type == DFGPU_FLOAT64, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: AVG takes a Float64 argument (the planner casts)");
      if ((ag.func == DFGPU_AGG_MIN || ag.func == DFGPU_AGG_MAX)) DF_CHECK(ag.arg_type != DFGPU_FLOAT32, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: MIN/MAX over Float32 stays on dfgpu_agg");
      if (ag.cls == C_DEC) DF_CHECK(ag.func == DFGPU_AGG_SUM || ag.func == DFGPU_AGG_COUNT, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: SUM / COUNT over Decimal128 (MIN / MAX / AVG stay on the CPU operator)");
      DF_CHECK(next < budget, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: not enough accumulator words in the lookup (n_acc_words)");
      ag.word = next++;
      if (ag.cls == C_DEC && ag.func == DFGPU_AGG_SUM) { DF_CHECK(next < budget, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: a Decimal128 SUM takes two accumulator words (n_acc_words)"); next++; }
      if (ag.func == DFGPU_AGG_AVG) { DF_CHECK(next < budget, DFGPU_ERR_UNSUPPORTED, "pipeline aggregate: not enough accumulator words in the lookup (n_acc_words)"); ag.cnt_word = next++; ag.nn_word = ag.cnt_word; }
    }
    new_aggs.push_back(std::move(ag));
  }
  // spare words become non-null counters (SUM / MIN / MAX of a nullable argument are NULL until a value arrives, accumulate.rs:164-188)
  for (auto& ag : new_aggs)
    if ((ag.func == DFGPU_AGG_SUM || ag.func == DFGPU_AGG_MIN || ag.func == DFGPU_AGG_MAX) && next < budget) ag.nn_word = next++;
  p->rows_word = rows_word;
  p->group_cols.assign(group_cols, group_cols + n_group);
  p->aggs = std::move(new_aggs);
  l->acc_claimed = true;
  p->agg_stage = stage; p->agg_mode = mode; p->batch_size = batch_size; p->sink = SINK_AGG;
  DF_API_END
}

int dfgpu_pipeline_sink_output(dfgpu_pipeline* p, const int32_t* out_cols, int32_t n_out, int64_t batch_size) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p && out_cols && n_out >= 1 && n_out <= kMaxPipeCols, DFGPU_ERR_INVALID, "pipeline output: 1..16 columns");
  DF_CHECK(p->sink == SINK_NONE && p->m_input_rows == 0, DFGPU_ERR_STATE, "pipeline: the sink is chosen once, before the first push");
  for (int c = 0; c < n_out; ++c) {
    DF_CHECK(out_cols[c] >= 0 && out_cols[c] < (int)p->vtypes.size(), DFGPU_ERR_INVALID, "pipeline output: column out of range");
    DF_CHECK(type_width(p->vtypes[out_cols[c]]) <= 8, DFGPU_ERR_UNSUPPORTED, "pipeline output: 16-byte columns leave through dfgpu_filter / dfgpu_hashjoin");
    p->out_cols.push_back(out_cols[c]);
  }
  p->batch_size = batch_size; p->sink = SINK_OUTPUT;
  DF_API_END
}

int dfgpu_pipeline_sink_output_unordered(dfgpu_pipeline* p, const int32_t* out_cols, int32_t n_out, int64_t batch_size) {
  const int rc = dfgpu_pipeline_sink_output(p, out_cols, n_out, batch_size);
  if (rc == DFGPU_OK) p->out_ordered = false;
  return rc;
}

int dfgpu_pipeline_set_name(dfgpu_pipeline* p, const char* name) {
  if (!p || !name) return DFGPU_ERR_INVALID;
  p->name = name;
  return DFGPU_OK;
}

int dfgpu_pipeline_push_device(dfgpu_pipeline* p, const dfgpu_column* cols, int32_t n_cols) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p && cols, DFGPU_ERR_INVALID, "null argument");
  std::vector<DCol> v;
  for (int i = 0; i < n_cols; ++i) v.push_back(device_view(cols[i]));
  pipeline_push(p, v);   // consumed (stream-synchronised) before returning
  DF_API_END
}
int dfgpu_pipeline_push_host(dfgpu_pipeline* p, const dfgpu_column* cols, int32_t n_cols) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p && cols, DFGPU_ERR_INVALID, "null argument");
  set_device(p->ctx);
  std::vector<DCol> v;
  for (int i = 0; i < n_cols; ++i) v.push_back(upload_column(p->ctx, cols[i]));
  pipeline_push(p, v);
  DF_API_END
}
int dfgpu_pipeline_push_arrow(dfgpu_pipeline* p, const struct ArrowArray* batch, const struct ArrowSchema* schema) {
  try {
    std::vector<dfgpu_column> cols = arrow_to_columns(batch, schema);
    return dfgpu_pipeline_push_host(p, cols.data(), (int32_t)cols.size());
  } catch (const Error& e) { if (p) p->ctx->last_error = e.what(); return e.code; }
}
int dfgpu_pipeline_finish(dfgpu_pipeline* p) {
  DF_API_BEGIN(p ? p->ctx : nullptr)
  DF_CHECK(p, DFGPU_ERR_INVALID, "null argument");
  pipeline_finish(p);
  DF_API_END
}
int dfgpu_pipeline_next(dfgpu_pipeline* p, int host, dfgpu_batch** out) {
  dfgpu_ctx* _ctx = p ? p->ctx : nullptr;
  try {
    DF_CHECK(p && out, DFGPU_ERR_INVALID, "null argument");
    if (p->outq.empty()) { *out = nullptr; return DFGPU_END; }
    BatchPtr b = std::move(p->outq.front());
    p->outq.pop_front();
    if (host) { set_device(p->ctx); b = to_host_batch(p->ctx, *b); }
    *out = b.release();
    return DFGPU_OK;
  } catch (const dfgpu::Error& e) { if (_ctx) _ctx->last_error = e.what(); return e.code; }
  catch (const std::exception& e) { if (_ctx) _ctx->last_error = e.what(); return DFGPU_ERR_INVALID; }
}
int64_t dfgpu_pipeline_metric(dfgpu_pipeline* p, const char* name) {
  if (!p || !name) return -1;
  std::string s(name);
  if (s == "input_rows") return p->m_input_rows;
  if (s == "sink_rows") return p->m_sink_rows;
  if (s == "output_rows") return p->m_output_rows;
  if (s == "num_groups") return p->m_groups;
  return -1;
}
void dfgpu_pipeline_destroy(dfgpu_pipeline* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  if (p->sink == SINK_AGG && p->agg_stage >= 0) p->stages[p->agg_stage].lookup->acc_claimed = false;
  delete p;
}

}  // extern "C"
