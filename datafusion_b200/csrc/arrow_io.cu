// arrow_io.cu — Arrow C Data Interface entry points: a record batch arrives as a struct array
// (`+s`) exactly as DataFusion's FFI ships it (reference datafusion/ffi/src/record_batch_stream.rs:101-110
// record_batch_to_wrapped_array; consumer side :151-167).  The caller keeps ownership of the
// ArrowArray; buffers are copied H2D (cudaMemcpyAsync on the ctx stream) before the call returns.
#include "batch.cuh"

namespace dfgpu {

static int arrow_type(const char* fmt) {
  if (!fmt) return -1;
  std::string f(fmt);
  if (f == "b") return DFGPU_BOOL;
  if (f == "c") return DFGPU_INT8;
  if (f == "C") return DFGPU_UINT8;
  if (f == "s") return DFGPU_INT16;
  if (f == "S") return DFGPU_UINT16;
  if (f == "i") return DFGPU_INT32;
  if (f == "I") return DFGPU_UINT32;
  if (f == "l") return DFGPU_INT64;
  if (f == "L") return DFGPU_UINT64;
  if (f == "f") return DFGPU_FLOAT32;
  if (f == "g") return DFGPU_FLOAT64;
  if (f == "tdD") return DFGPU_DATE32;
  if (f == "tdm") return DFGPU_DATE64;
  if (f.rfind("ts", 0) == 0) return DFGPU_TIMESTAMP;
  if (f.rfind("d:", 0) == 0) {
    // decimal128 only ("d:p,s" or "d:p,s,128")
    int commas = 0;
    for (char ch : f) commas += ch == ',';
    if (commas == 1 || f.size() >= 4 && f.substr(f.size() - 4) == ",128") {
      int p = 0, sc = 0;
      if (sscanf(f.c_str(), "d:%d,%d", &p, &sc) == 2 && p >= 1 && p <= 38 && sc >= 0 && sc <= p) return dec_type(p, sc);
      return DFGPU_DECIMAL128;  // negative scale: carried as an opaque 16-byte value
    }
    return -1;
  }
  return -1;
}

std::vector<dfgpu_column> arrow_to_columns(const ArrowArray* batch, const ArrowSchema* schema) {
  DF_CHECK(batch && schema, DFGPU_ERR_INVALID, "null Arrow pointers");
  DF_CHECK(schema->format && std::string(schema->format) == "+s", DFGPU_ERR_INVALID, "expected a struct array (record batch) at the top level");
  DF_CHECK(batch->n_children == schema->n_children, DFGPU_ERR_INVALID, "Arrow array/schema children mismatch");
  DF_CHECK(batch->null_count <= 0, DFGPU_ERR_UNSUPPORTED, "top-level struct nulls are not supported");
  std::vector<dfgpu_column> cols;
  for (int64_t i = 0; i < batch->n_children; ++i) {
    const ArrowArray* a = batch->children[i];
    const ArrowSchema* s = schema->children[i];
    int t = arrow_type(s->format);
    DF_CHECK(t > 0, DFGPU_ERR_UNSUPPORTED, std::string("Arrow format not supported on the GPU path: ") + (s->format ? s->format : "?"));
    DF_CHECK(a->dictionary == nullptr, DFGPU_ERR_UNSUPPORTED, "dictionary arrays are not supported on the GPU path");
    DF_CHECK(a->n_buffers == 2, DFGPU_ERR_INVALID, "primitive Arrow array must have 2 buffers");
    dfgpu_column c;
    memset(&c, 0, sizeof(c));
    c.type = t;
    // a sliced struct applies its offset/length to the children
    c.length = batch->length;
    c.offset = a->offset + batch->offset;
    DF_CHECK(a->length >= batch->offset + batch->length - 0 || a->length == batch->length, DFGPU_ERR_INVALID, "child shorter than the record batch");
    c.null_count = (batch->offset == 0 && a->length == batch->length) ? a->null_count : -1;
    c.validity = (const uint8_t*)a->buffers[0];
    c.values = a->buffers[1];
    if (!c.validity) c.null_count = 0;
    cols.push_back(c);
  }
  return cols;
}

}  // namespace dfgpu

using namespace dfgpu;

extern "C" {

int dfgpu_filter_push_arrow(dfgpu_filter* f, const struct ArrowArray* batch, const struct ArrowSchema* schema) {
  try {
    std::vector<dfgpu_column> cols = arrow_to_columns(batch, schema);
    return dfgpu_filter_push_host(f, cols.data(), (int32_t)cols.size());
  } catch (const Error& e) { return e.code; }
}
int dfgpu_hashjoin_push_build_arrow(dfgpu_hashjoin* j, const struct ArrowArray* batch, const struct ArrowSchema* schema) {
  try {
    std::vector<dfgpu_column> cols = arrow_to_columns(batch, schema);
    return dfgpu_hashjoin_push_build_host(j, cols.data(), (int32_t)cols.size());
  } catch (const Error& e) { return e.code; }
}
int dfgpu_hashjoin_push_probe_arrow(dfgpu_hashjoin* j, const struct ArrowArray* batch, const struct ArrowSchema* schema) {
  try {
    std::vector<dfgpu_column> cols = arrow_to_columns(batch, schema);
    return dfgpu_hashjoin_push_probe_host(j, cols.data(), (int32_t)cols.size());
  } catch (const Error& e) { return e.code; }
}
int dfgpu_agg_push_arrow(dfgpu_agg* a, const struct ArrowArray* batch, const struct ArrowSchema* schema) {
  try {
    std::vector<dfgpu_column> cols = arrow_to_columns(batch, schema);
    return dfgpu_agg_push_host(a, cols.data(), (int32_t)cols.size());
  } catch (const Error& e) { return e.code; }
}

}  // extern "C"
