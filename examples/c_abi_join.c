/* examples/c_abi_join.c — a plain-C consumer of include/dfgpu.h: the join_inner_one fixture of the reference
 * (datafusion/physical-plan/src/joins/hash_join/exec.rs:3311-3360: build (a1,b1,c1), probe (a2,b1,c2), Inner on b1)
 * pushed through the C ABI with host buffers.  Shows what the Rust shim's FFI calls look like, proves the header is
 * valid C and the library links without any C++/torch/Python dependency:
 *     gcc -I include examples/c_abi_join.c -L datafusion_b200 -ldfgpu -Wl,-rpath,$PWD/datafusion_b200 -o c_abi_join
 * Exit codes: 0 = rows printed and equal to the reference snapshot, 2 = no CUDA device (there is no CPU fallback), 1 = error. */
#include <stdio.h>
#include <string.h>
#include "dfgpu.h"

static dfgpu_column col_i32(const int32_t* v, int64_t n) {
  dfgpu_column c;
  memset(&c, 0, sizeof(c));
  c.type = DFGPU_INT32; c.length = n; c.null_count = 0; c.values = v; c.validity = NULL;
  return c;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ < 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ctx ? dfgpu_last_error(ctx) : "?"); return 1; } } while (0)

int main(void) {
  dfgpu_ctx* ctx = NULL;
  if (dfgpu_device_count() == 0 || dfgpu_ctx_create(0, NULL, &ctx) < 0) {
    fprintf(stderr, "no CUDA device: libdfgpu has no CPU fallback\n");
    return 2;
  }
  const int32_t a1[] = {1, 2, 3}, b1[] = {4, 5, 5}, c1[] = {7, 8, 9};        /* build side; b1 = 5 is repeated */
  const int32_t a2[] = {10, 20, 30}, b2[] = {4, 5, 6}, c2[] = {70, 80, 90};  /* probe side */
  const int32_t types[3] = {DFGPU_INT32, DFGPU_INT32, DFGPU_INT32};
  const int32_t on_build[1] = {1}, on_probe[1] = {1};
  const int32_t out_side[6] = {0, 0, 0, 1, 1, 1}, out_index[6] = {0, 1, 2, 0, 1, 2};
  dfgpu_hashjoin_options opt;
  dfgpu_hashjoin_default_options(&opt);          /* Inner, NullEqualsNothing, batch_size 8192, perfect-hash thresholds of config.rs */
  dfgpu_hashjoin* j = NULL;
  CHECK(dfgpu_hashjoin_create(ctx, types, 3, types, 3, on_build, on_probe, 1, out_side, out_index, 6, &opt, &j));
  dfgpu_column build[3] = {col_i32(a1, 3), col_i32(b1, 3), col_i32(c1, 3)};
  dfgpu_column probe[3] = {col_i32(a2, 3), col_i32(b2, 3), col_i32(c2, 3)};
  CHECK(dfgpu_hashjoin_push_build_host(j, build, 3));
  CHECK(dfgpu_hashjoin_finish_build(j));
  CHECK(dfgpu_hashjoin_push_probe_host(j, probe, 3));
  CHECK(dfgpu_hashjoin_finish_probe(j));
  /* the reference's snapshot (sorted): (1,4,7,10,4,70) (2,5,8,20,5,80) (3,5,9,20,5,80); emission order = probe order x ascending build row */
  const int32_t expected[3][6] = {{1, 4, 7, 10, 4, 70}, {2, 5, 8, 20, 5, 80}, {3, 5, 9, 20, 5, 80}};
  int64_t seen = 0;
  int ok = 1;
  for (;;) {
    dfgpu_batch* b = NULL;
    int rc = dfgpu_hashjoin_next(j, /*host=*/1, &b);
    if (rc == DFGPU_END) break;
    CHECK(rc);
    const int64_t rows = dfgpu_batch_num_rows(b);
    dfgpu_column out[6];
    for (int c = 0; c < 6; ++c) CHECK(dfgpu_batch_column(b, c, &out[c]));
    for (int64_t r = 0; r < rows; ++r, ++seen) {
      printf("row %lld:", (long long)seen);
      for (int c = 0; c < 6; ++c) {
        const int32_t v = ((const int32_t*)out[c].values)[r];
        printf(" %d", v);
        if (seen < 3 && v != expected[seen][c]) ok = 0;
      }
      printf("\n");
    }
    dfgpu_batch_release(b);
  }
  dfgpu_hashjoin_destroy(j);
  dfgpu_ctx_destroy(ctx);
  if (seen != 3 || !ok) { fprintf(stderr, "unexpected join output\n"); return 1; }
  printf("join_inner_one: 3 rows, equal to the reference snapshot\n");
  return 0;
}
