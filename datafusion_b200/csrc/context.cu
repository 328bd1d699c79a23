// context.cu — ctx / memory / timing / generators / batch accessors of the C ABI (include/dfgpu.h).
#include "batch.cuh"

using namespace dfgpu;

namespace dfgpu {

// deterministic synthetic data: the same functions exist in oracle/gen.h so CPU and
// GPU inputs are bit-identical without ever crossing PCIe.
__host__ __device__ inline uint64_t perm_bijection(uint64_t i, uint64_t n, uint64_t seed) {
  // cycle-walking Feistel permutation over [0, n)
  int bits = 1;
  while ((1ull << bits) < n) ++bits;
  if (bits & 1) ++bits;
  const int half = bits / 2;
  const uint64_t mask = (1ull << half) - 1ull;
  uint64_t x = i;
  do {
    uint64_t l = x >> half, r = x & mask;
    for (int round = 0; round < 4; ++round) {
      uint64_t f = mix64(r + seed * 0x9E3779B97F4A7C15ull + (uint64_t)round * 0xD1B54A32D192ED03ull) & mask;
      uint64_t nl = r, nr = l ^ f;
      l = nl; r = nr;
    }
    x = (l << half) | r;
  } while (x >= n);
  return x;
}

__global__ void generate_i64_kernel(int kind, uint64_t seed, int64_t a, int64_t b, int64_t start, int64_t n, int64_t* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t idx = (uint64_t)(start + i);
    int64_t v;
    switch (kind) {
      case DFGPU_GEN_SEQ: v = a + (int64_t)idx; break;
      case DFGPU_GEN_UNIFORM: v = a + (int64_t)(splitmix64_at(seed, idx) % (uint64_t)b); break;
      case DFGPU_GEN_SPLITMIX: v = (int64_t)splitmix64_at(seed, idx); break;
      case DFGPU_GEN_PERM: v = a + (int64_t)perm_bijection(idx, (uint64_t)b, seed); break;
      case DFGPU_GEN_SPARSE_OF: v = (int64_t)splitmix64_at(seed, splitmix64_at((uint64_t)a, idx) % (uint64_t)b); break;
      default: v = 0;
    }
    out[i] = v;
  }
}

__global__ void l2_flush_kernel(int4* p, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_int4((int)i, 1, 2, 3);
}

}  // namespace dfgpu

extern "C" {

const char* dfgpu_version(void) { return "dfgpu 0.1.0 (sm_100a)"; }

int dfgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int dfgpu_ctx_create(int device, void* stream, dfgpu_ctx** out) {
  if (!out) return DFGPU_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); return DFGPU_ERR_CUDA; }
  std::unique_ptr<dfgpu_ctx> ctx(new dfgpu_ctx());
  try {
    ctx->device = device;
    DF_CUDA(cudaSetDevice(device));
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else { DF_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    DF_CUDA(cudaMallocHost(&ctx->pinned_scalar, 256));
    // idle-block cache limit: half of the device memory (SF100-sized pipelines free > 24 GB of temporaries per step; trimming
    // them costs a device sync + cudaFree / cudaMalloc round trips every step), DFGPU_DEV_CACHE_GB overrides
    {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) ctx->dev_cache_limit = std::max<size_t>(ctx->dev_cache_limit, total_b / 2);
      else cudaGetLastError();
      if (const char* e = getenv("DFGPU_DEV_CACHE_GB")) { const long gb = atol(e); if (gb > 0) ctx->dev_cache_limit = (size_t)gb << 30; }
    }
    // keep freed blocks in the pool: operators re-allocate similar sizes every batch
    cudaMemPool_t pool;
    DF_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    DF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  } catch (const Error&) { return DFGPU_ERR_CUDA; }
  *out = ctx.release();
  return DFGPU_OK;
}

void dfgpu_ctx_destroy(dfgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  dev_cache_trim(ctx);
  if (ctx->l2_scratch) cudaFree(ctx->l2_scratch);
  if (ctx->pinned_scalar) cudaFreeHost(ctx->pinned_scalar);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* dfgpu_last_error(dfgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }
void* dfgpu_ctx_stream(dfgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t dfgpu_launch_count(dfgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dfgpu_set_kernel_timing(dfgpu_ctx* ctx, int enabled) {
  if (!ctx) return DFGPU_ERR_INVALID;
  ctx->time_kernels = enabled != 0;
  return DFGPU_OK;
}
int dfgpu_kernel_time(dfgpu_ctx* ctx, const char* name, double* total_ms, int64_t* count) {
  DF_API_BEGIN(ctx)
  DF_CHECK(name && total_ms && count, DFGPU_ERR_INVALID, "null argument");
  *total_ms = 0; *count = 0;
  for (auto& t : ctx->timings) {
    if (t.name != name) continue;
    for (auto& pr : t.pending) {
      DF_CUDA(cudaEventSynchronize(pr.second));
      float ms = 0;
      DF_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
      t.total_ms += ms; t.count++;
      cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    t.pending.clear();
    *total_ms = t.total_ms; *count = t.count;
  }
  DF_API_END
}
int dfgpu_kernel_time_reset(dfgpu_ctx* ctx) {
  DF_API_BEGIN(ctx)
  for (auto& t : ctx->timings) {
    for (auto& pr : t.pending) { cudaEventSynchronize(pr.second); cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    t.pending.clear(); t.total_ms = 0; t.count = 0;
  }
  DF_API_END
}

int dfgpu_sync(dfgpu_ctx* ctx) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  DF_CUDA(cudaStreamSynchronize(ctx->stream));
  DF_API_END
}
// non-blocking: 1 = every piece of work queued on the ctx stream has completed, 0 = still running, < 0 = error.  The waker side of
// the async contract: a Gpu*Exec stream's poll_next returns Poll::Pending while this is 0 instead of blocking a tokio worker
// (execution_plan.rs:549-563 "must yield regularly").
int dfgpu_poll_ready(dfgpu_ctx* ctx) {
  if (!ctx) return DFGPU_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaStreamQuery(ctx->stream);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  ctx->last_error = std::string("CUDA error in dfgpu_poll_ready: ") + cudaGetErrorString(e);
  return DFGPU_ERR_CUDA;
}
int dfgpu_malloc(dfgpu_ctx* ctx, size_t bytes, void** out) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  *out = dev_alloc(ctx, bytes ? bytes : 8);
  DF_API_END
}
int dfgpu_free(dfgpu_ctx* ctx, void* p) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  if (p) dev_free(ctx, p);
  DF_API_END
}
int dfgpu_host_alloc(dfgpu_ctx* ctx, size_t bytes, void** out) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  DF_CUDA(cudaMallocHost(out, bytes ? bytes : 8));
  DF_API_END
}
int dfgpu_host_free(dfgpu_ctx* ctx, void* p) {
  DF_API_BEGIN(ctx)
  if (p) DF_CUDA(cudaFreeHost(p));
  DF_API_END
}
// Page-lock caller-owned host memory in place (an Arrow buffer the Rust side already holds): later `*_push_host` / `*_push_arrow`
// calls over it copy at the pinned PCIe rate instead of through the driver's pageable staging path.
int dfgpu_host_register(dfgpu_ctx* ctx, void* p, size_t bytes) {
  DF_API_BEGIN(ctx)
  DF_CHECK(p && bytes, DFGPU_ERR_INVALID, "null argument");
  set_device(ctx);
  DF_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  DF_API_END
}
int dfgpu_host_unregister(dfgpu_ctx* ctx, void* p) {
  DF_API_BEGIN(ctx)
  if (p) DF_CUDA(cudaHostUnregister(p));
  DF_API_END
}
int dfgpu_memcpy_h2d(dfgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  if (bytes) DF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  DF_API_END
}
int dfgpu_memcpy_d2h(dfgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  if (bytes) DF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DF_API_END
}
int dfgpu_memset(dfgpu_ctx* ctx, void* dst, int value, size_t bytes) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  if (bytes) DF_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
  DF_API_END
}
int dfgpu_trim_device_cache(dfgpu_ctx* ctx) {
  DF_API_BEGIN(ctx)
  DF_CHECK(ctx, DFGPU_ERR_INVALID, "null ctx");
  set_device(ctx);
  dev_cache_trim(ctx);
  DF_API_END
}

int dfgpu_flush_l2(dfgpu_ctx* ctx) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  if (!ctx->l2_scratch) {
    ctx->l2_scratch_bytes = 256ull << 20;  // 2x the 126 MB L2
    DF_CUDA(cudaMalloc(&ctx->l2_scratch, ctx->l2_scratch_bytes));
  }
  l2_flush_kernel<<<kNumSMs * 8, 256, 0, ctx->stream>>>((int4*)ctx->l2_scratch, ctx->l2_scratch_bytes / 16);
  DF_CUDA(cudaGetLastError());  // not counted in launches: bench hygiene, not product work
  DF_API_END
}

int dfgpu_event_create(dfgpu_ctx* ctx, void** out) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  cudaEvent_t ev;
  DF_CUDA(cudaEventCreate(&ev));
  *out = (void*)ev;
  DF_API_END
}
int dfgpu_event_record(dfgpu_ctx* ctx, void* ev) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  DF_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
  DF_API_END
}
int dfgpu_event_elapsed_ms(dfgpu_ctx* ctx, void* start, void* stop, float* ms) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  DF_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  DF_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  DF_API_END
}
int dfgpu_event_destroy(dfgpu_ctx* ctx, void* ev) {
  DF_API_BEGIN(ctx)
  if (ev) DF_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  DF_API_END
}

int dfgpu_generate_i64(dfgpu_ctx* ctx, int kind, uint64_t seed, int64_t a, int64_t b, int64_t start, int64_t n, int64_t* out_device) {
  DF_API_BEGIN(ctx)
  set_device(ctx);
  DF_CHECK(kind >= 0 && kind <= DFGPU_GEN_SPARSE_OF, DFGPU_ERR_INVALID, "unknown generator kind");
  if ((kind == DFGPU_GEN_UNIFORM || kind == DFGPU_GEN_PERM || kind == DFGPU_GEN_SPARSE_OF)) DF_CHECK(b > 0, DFGPU_ERR_INVALID, "generator range must be positive");
  if (n > 0) {
    generate_i64_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, ctx->stream>>>(kind, seed, a, b, start, n, out_device);
    DF_CUDA(cudaGetLastError());
  }
  DF_API_END
}

// ---- output batches ----
int64_t dfgpu_batch_num_rows(const dfgpu_batch* b) { return b ? b->rows : -1; }
int32_t dfgpu_batch_num_columns(const dfgpu_batch* b) { return b ? (int32_t)(b->host ? b->hcols.size() : b->cols.size()) : -1; }
int dfgpu_batch_is_host(const dfgpu_batch* b) { return b && b->host ? 1 : 0; }

int dfgpu_batch_column(const dfgpu_batch* b, int32_t i, dfgpu_column* out) {
  if (!b || !out || i < 0 || i >= dfgpu_batch_num_columns(b)) return DFGPU_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  if (b->host) {
    const HCol& h = b->hcols[i];
    out->type = h.type; out->length = h.length; out->offset = 0; out->null_count = h.null_count;
    out->values = h.values ? h.values->ptr : nullptr;
    out->validity = h.validity ? (const uint8_t*)h.validity->ptr : nullptr;
  } else {
    const DCol& d = b->cols[i];
    out->type = d.type; out->length = d.length; out->null_count = d.null_count;
    out->values = d.values; out->validity = d.validity;
    // DCol keeps `values` pre-advanced for fixed-width types; report offset 0 unless a bitmap needs it
    out->offset = (d.type == DFGPU_BOOL || d.validity) ? d.offset : 0;
    if (d.type != DFGPU_BOOL && d.validity && d.offset != 0) {
      // express as an Arrow-style slice: step the values pointer back by the validity offset
      out->values = (const char*)d.values - d.offset * type_width(d.type);
    }
  }
  return DFGPU_OK;
}

void dfgpu_batch_release(dfgpu_batch* b) {
  if (!b) return;
  if (b->ctx) cudaSetDevice(b->ctx->device);
  delete b;
}

// ---- Arrow C Data export of a host batch (struct array, one child per column) ----
namespace {
const char* arrow_format(int type) {
  switch (type & 0xff) {
    case DFGPU_BOOL: return "b";
    case DFGPU_INT8: return "c"; case DFGPU_UINT8: return "C";
    case DFGPU_INT16: return "s"; case DFGPU_UINT16: return "S";
    case DFGPU_INT32: return "i"; case DFGPU_UINT32: return "I";
    case DFGPU_INT64: return "l"; case DFGPU_UINT64: return "L";
    case DFGPU_FLOAT32: return "f"; case DFGPU_FLOAT64: return "g";
    case DFGPU_DATE32: return "tdD"; case DFGPU_DATE64: return "tdm";
    case DFGPU_TIMESTAMP: return "tsn:";
    case DFGPU_DECIMAL128: return "d:38,0";
    default: return "n";
  }
}
// Arrow C Data Interface ownership: every child is an independently released structure (a consumer may MOVE a child out of the
// parent — copy the struct, null the original's release — and keep it after the parent is released), the parent owns the child
// struct memory and releases the children that were not moved.
struct ChildPrivate { HCol col; const void* buffers[2]; };          // keeps this column's pinned buffers alive
struct ExportPrivate { std::vector<ArrowArray*> children; const void* top_buffers[1] = {nullptr}; };
struct ChildSchemaPrivate { std::string name, format; };
struct SchemaPrivate { std::vector<ArrowSchema*> children; };
void release_child_array(ArrowArray* a) {
  if (!a || !a->release) return;
  delete (ChildPrivate*)a->private_data;
  a->release = nullptr;
}
void release_top_array(ArrowArray* a) {
  if (!a || !a->release) return;
  auto* p = (ExportPrivate*)a->private_data;
  for (ArrowArray* c : p->children) { if (c->release) c->release(c); delete c; }
  delete p;
  a->release = nullptr;
}
void release_child_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  delete (ChildSchemaPrivate*)s->private_data;
  s->release = nullptr;
}
void release_top_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  auto* p = (SchemaPrivate*)s->private_data;
  for (ArrowSchema* c : p->children) { if (c->release) c->release(c); delete c; }
  delete p;
  s->release = nullptr;
}
}  // namespace

int dfgpu_batch_export_arrow(dfgpu_batch* b, struct ArrowArray* out_array, struct ArrowSchema* out_schema) {
  if (!b || !b->host || !out_array || !out_schema) return DFGPU_ERR_INVALID;
  auto* p = new ExportPrivate();
  const size_t n = b->hcols.size();
  for (size_t i = 0; i < n; ++i) {
    auto* cp = new ChildPrivate();
    cp->col = b->hcols[i];  // shared_ptr copies: the buffers outlive the dfgpu_batch and the parent array
    cp->buffers[0] = cp->col.validity ? cp->col.validity->ptr : nullptr;
    cp->buffers[1] = cp->col.values ? cp->col.values->ptr : nullptr;
    auto* a = new ArrowArray();
    memset(a, 0, sizeof(*a));
    a->length = cp->col.length; a->null_count = cp->col.validity ? cp->col.null_count : 0; a->offset = 0;
    a->n_buffers = 2; a->buffers = cp->buffers;
    a->release = release_child_array; a->private_data = cp;
    p->children.push_back(a);
  }
  memset(out_array, 0, sizeof(*out_array));
  out_array->length = b->rows; out_array->null_count = 0; out_array->offset = 0;
  out_array->n_buffers = 1; out_array->buffers = p->top_buffers;
  out_array->n_children = (int64_t)n; out_array->children = p->children.data();
  out_array->release = release_top_array; out_array->private_data = p;

  auto* sp = new SchemaPrivate();
  for (size_t i = 0; i < n; ++i) {
    auto* csp = new ChildSchemaPrivate();
    csp->name = "c" + std::to_string(i);
    const int t = b->hcols[i].type;
    csp->format = arrow_format(t);
    if (type_is_decimal(t) && dec_precision(t) > 0) csp->format = "d:" + std::to_string(dec_precision(t)) + "," + std::to_string(dec_scale(t));   // "d:precision,scale"
    auto* cs = new ArrowSchema();
    memset(cs, 0, sizeof(*cs));
    cs->format = csp->format.c_str(); cs->name = csp->name.c_str(); cs->flags = ARROW_FLAG_NULLABLE;
    cs->release = release_child_schema; cs->private_data = csp;
    sp->children.push_back(cs);
  }
  memset(out_schema, 0, sizeof(*out_schema));
  out_schema->format = "+s"; out_schema->name = ""; out_schema->n_children = (int64_t)n;
  out_schema->children = sp->children.data(); out_schema->release = release_top_schema; out_schema->private_data = sp;
  return DFGPU_OK;
}

}  // extern "C"
