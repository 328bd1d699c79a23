// dictionary.cu — dictionary-coded string keys for the GPU operators.
//
// The reference joins and groups on Utf8 / Utf8View / Dictionary(_, Utf8) keys by hashing and comparing the bytes
// (hash_utils.rs create_hashes, group_values/mod.rs:139-217 picks GroupValuesBytes / GroupValuesBytesView).  On the GPU path a
// string key is an INT32 code: equal strings <=> equal codes, so every integer-key kernel (hash join, group-by, filter
// `col = 'literal'`, exchange) applies unchanged.  That only holds inside ONE code space, while Arrow dictionaries are per
// batch: this object unifies them on the host (a hash map of the distinct strings — small next to the row count, which is
// why it is dictionary-coded in the first place) and a gather kernel rewrites the codes of a batch on the device.
#include "batch.cuh"
#include <string>
#include <unordered_map>

using namespace dfgpu;

struct dfgpu_dictionary {
  dfgpu_ctx* ctx = nullptr;
  std::unordered_map<std::string, int32_t> code_of;
  std::vector<std::string> values;
};

namespace dfgpu {

// out[i] = remap[codes[i]] for valid rows; one warp writes the 32 validity bits of its rows
template <typename T>
__global__ void __launch_bounds__(256) dict_remap_kernel(const T* __restrict__ codes, const uint8_t* __restrict__ valid, int64_t voff, int64_t n,
                                                      const int32_t* __restrict__ remap, int64_t n_remap, int32_t* __restrict__ out,
                                                      uint32_t* __restrict__ out_valid, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (n + 31) / 32;
  for (int64_t wi = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; wi < nw; wi += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t i = wi * 32 + lane;
    bool ok = false;
    if (i < n) {
      ok = !(valid && !bit_get(valid, voff + i));
      int32_t v = 0;
      if (ok) {
        const int64_t c = (int64_t)codes[i];
        if (c < 0 || c >= n_remap) { *bad = 1; ok = false; }
        else {
          v = remap[c];
          if (v < 0) ok = false;   // a NULL dictionary value: the row is NULL (DictionaryArray logical nulls)
        }
      }
      out[i] = ok ? v : 0;
    }
    const uint32_t b = __ballot_sync(0xffffffffu, ok);
    if (lane == 0 && out_valid) out_valid[wi] = b;
  }
}

}  // namespace dfgpu

extern "C" {

int dfgpu_dictionary_create(dfgpu_ctx* ctx, dfgpu_dictionary** out) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx && out, DFGPU_ERR_INVALID, "null argument");
  auto* d = new dfgpu_dictionary();
  d->ctx = ctx;
  *out = d;
  DF_API_END
}

int dfgpu_dictionary_unify(dfgpu_dictionary* d, const int32_t* offsets, const uint8_t* data, const uint8_t* validity, int64_t n_values, int32_t* remap_out) {
  DF_API_BEGIN(d ? d->ctx : nullptr)
  DF_CHECK(d && (n_values == 0 || (offsets && remap_out)), DFGPU_ERR_INVALID, "null argument");
  for (int64_t i = 0; i < n_values; ++i) {
    if (validity && !((validity[i >> 3] >> (i & 7)) & 1)) { remap_out[i] = -1; continue; }
    DF_CHECK(offsets[i + 1] >= offsets[i], DFGPU_ERR_INVALID, "dictionary: offsets must not decrease");
    std::string s(data ? (const char*)data + offsets[i] : "", (size_t)(offsets[i + 1] - offsets[i]));
    auto it = d->code_of.find(s);
    if (it == d->code_of.end()) {
      DF_CHECK(d->values.size() < (size_t)INT32_MAX, DFGPU_ERR_UNSUPPORTED, "dictionary: more than 2^31 distinct values");
      const int32_t code = (int32_t)d->values.size();
      d->values.push_back(s);
      it = d->code_of.emplace(std::move(s), code).first;
    }
    remap_out[i] = it->second;
  }
  DF_API_END
}

int32_t dfgpu_dictionary_code(dfgpu_dictionary* d, const uint8_t* bytes, int64_t len) {
  if (!d || len < 0) return -1;
  auto it = d->code_of.find(std::string((const char*)bytes, (size_t)len));
  return it == d->code_of.end() ? -1 : it->second;
}

int64_t dfgpu_dictionary_size(dfgpu_dictionary* d) { return d ? (int64_t)d->values.size() : 0; }

int dfgpu_dictionary_value(dfgpu_dictionary* d, int32_t code, const uint8_t** bytes, int64_t* len) {
  if (!d || !bytes || !len || code < 0 || (size_t)code >= d->values.size()) return DFGPU_ERR_INVALID;
  *bytes = (const uint8_t*)d->values[code].data();
  *len = (int64_t)d->values[code].size();
  return DFGPU_OK;
}

int dfgpu_dictionary_remap(dfgpu_dictionary* d, const dfgpu_column* codes, int codes_on_host, const int32_t* remap, int64_t n_remap, dfgpu_batch** out) {
  DF_API_BEGIN(d ? d->ctx : nullptr)
  DF_CHECK(d && codes && out && (remap || n_remap == 0), DFGPU_ERR_INVALID, "null argument");
  dfgpu_ctx* ctx = d->ctx;
  set_device(ctx);
  DF_CHECK(type_is_int(codes->type), DFGPU_ERR_INVALID, "dictionary: the codes column must be an integer column");
  DCol c = codes_on_host ? upload_column(ctx, *codes) : device_view(*codes);
  const int64_t n = c.length;
  DevBuf rm(ctx, (size_t)std::max<int64_t>(n_remap, 1) * 4), bad(ctx, 4);
  if (n_remap) DF_CUDA(cudaMemcpyAsync(rm.ptr, remap, (size_t)n_remap * 4, cudaMemcpyHostToDevice, ctx->stream));
  bad.zero();
  DCol o = alloc_col(ctx, DFGPU_INT32, n, true);
  if (n > 0) {
    const int grid = grid_for(n, 256, kNumSMs * 8);
    int32_t* op = (int32_t*)o.own_values->ptr;
    uint32_t* ov = o.own_validity->as<uint32_t>();
#define DF_REMAP(T) dict_remap_kernel<T><<<grid, 256, 0, ctx->stream>>>((const T*)c.values, c.validity, c.offset, n, rm.as<int32_t>(), n_remap, op, ov, bad.as<int>())
    switch (c.type) {
      case DFGPU_INT8: DF_REMAP(int8_t); break;
      case DFGPU_INT16: DF_REMAP(int16_t); break;
      case DFGPU_INT32: case DFGPU_DATE32: DF_REMAP(int32_t); break;
      case DFGPU_UINT8: DF_REMAP(uint8_t); break;
      case DFGPU_UINT16: DF_REMAP(uint16_t); break;
      case DFGPU_UINT32: DF_REMAP(uint32_t); break;
      default: DF_REMAP(int64_t); break;
    }
#undef DF_REMAP
    DF_LAUNCH_CHECK(ctx);
  }
  int hbad = 0;
  DF_CUDA(cudaMemcpyAsync(&hbad, bad.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
  DF_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller's remap table and codes are consumed before returning
  DF_CHECK(!hbad, DFGPU_ERR_INVALID, "dictionary: a code lies outside the batch's dictionary");
  o.null_count = -1;
  BatchPtr b(new dfgpu_batch());
  b->ctx = ctx; b->rows = n; b->host = false;
  b->cols.push_back(std::move(o));
  *out = b.release();
  DF_API_END
}

void dfgpu_dictionary_destroy(dfgpu_dictionary* d) { delete d; }

}  // extern "C"
