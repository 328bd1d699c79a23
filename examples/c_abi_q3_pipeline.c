/* examples/c_abi_q3_pipeline.c — the TPC-H Q3 physical plan (sqllogictest/test_files/tpch/plans/q3.slt.part:60-76) driven through the
 * C ABI as three fused pipelines, with the reference's money types (benchmarks/src/tpch/mod.rs:52-122: l_extendedprice, l_discount
 * Decimal128(15,2); sum(l_extendedprice * (1 - l_discount)) is Decimal128(38,4)):
 *
 *   P1  customer : FilterExec(c_mktsegment = 1)                      -> build L1 = key set {c_custkey}
 *   P2  orders   : FilterExec(o_orderdate < CUT) -> RightSemi vs L1  -> build L2 = {o_orderkey -> (o_orderdate, o_shippriority)}
 *   P3  lineitem : FilterExec(l_shipdate > CUT)  -> Inner vs L2      -> AggregateExec gby [l_orderkey, o_orderdate, o_shippriority]
 *
 * What a Rust GpuPipelineExec calls, in plain C99.  The program checks the rows against a nested-loop evaluation
 * of the same tiny tables.
 *     gcc -std=c99 -I include examples/c_abi_q3_pipeline.c -L datafusion_b200 -ldfgpu -Wl,-rpath,$PWD/datafusion_b200 -o c_abi_q3
 * Exit codes: 0 = rows equal to the nested-loop evaluation, 2 = no CUDA device (there is no CPU fallback), 1 = error or mismatch. */
#include <stdio.h>
#include <string.h>
#include "dfgpu.h"

#define NC 6
#define NO 8
#define NL 14
#define CUT 9204

static dfgpu_column col(int32_t type, const void* v, int64_t n) {
  dfgpu_column c;
  memset(&c, 0, sizeof(c));
  c.type = type; c.length = n; c.null_count = 0; c.values = v; c.validity = NULL;
  return c;
}
static dfgpu_expr_node node(int32_t kind, int32_t a, int32_t type, int64_t lit) {
  dfgpu_expr_node n;
  memset(&n, 0, sizeof(n));
  n.kind = kind; n.a = a; n.type = type; n.lit_i64 = lit;   /* Decimal128 literal: low 64 bits here, the high 64 bits (0) in lit_f64's bytes */
  return n;
}

#define CHECK(call) do { int rc_ = (call); if (rc_ < 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ctx ? dfgpu_last_error(ctx) : "?"); return 1; } } while (0)

int main(void) {
  dfgpu_ctx* ctx = NULL;
  if (dfgpu_device_count() == 0 || dfgpu_ctx_create(0, NULL, &ctx) < 0) {
    fprintf(stderr, "no CUDA device: libdfgpu has no CPU fallback\n");
    return 2;
  }
  /* ---- tables ---- */
  const int64_t c_custkey[NC] = {1, 2, 3, 4, 5, 6}, c_mktsegment[NC] = {1, 0, 1, 3, 1, 2};
  const int64_t o_orderkey[NO] = {100, 101, 102, 103, 104, 105, 106, 107}, o_custkey[NO] = {1, 2, 3, 3, 5, 6, 1, 4};
  const int32_t o_orderdate[NO] = {9000, 9100, 9300, 9150, 9203, 9000, 9204, 9100}, o_shippriority[NO] = {0, 0, 1, 0, 1, 0, 0, 1};
  const int64_t l_orderkey[NL] = {100, 100, 101, 102, 103, 103, 103, 104, 105, 106, 999, 100, 104, 103};
  /* Decimal128(15,2): 16-byte little-endian two's complement = {low word, high word} */
  uint64_t l_extendedprice[NL][2], l_discount[NL][2];
  const int64_t price_cents[NL] = {1000000, 250050, 99999, 123456, 5000000, 770000, 31415, 2718281, 100, 4242, 77, 999999999, 1, 1050};
  const int64_t disc_pct[NL] = {5, 0, 10, 7, 4, 10, 0, 3, 1, 2, 9, 6, 10, 8};
  const int32_t l_shipdate[NL] = {9300, 9100, 9300, 9400, 9205, 9204, 9999, 9250, 9300, 9300, 9300, 9210, 9203, 9206};
  for (int i = 0; i < NL; ++i) { l_extendedprice[i][0] = (uint64_t)price_cents[i]; l_extendedprice[i][1] = 0; l_discount[i][0] = (uint64_t)disc_pct[i]; l_discount[i][1] = 0; }
  const int32_t money = DFGPU_DECIMAL128_TYPE(15, 2);

  /* ---- P1: customer -> key set ---- */
  dfgpu_lookup_options lo;
  dfgpu_lookup_default_options(&lo);
  lo.has_key_range = 1; lo.key_min = 1; lo.key_max = NC;            /* column statistics: a dense range becomes a bitmap */
  dfgpu_lookup* l1 = NULL;
  CHECK(dfgpu_lookup_create(ctx, DFGPU_INT64, NULL, 0, &lo, &l1));
  const int32_t c_types[2] = {DFGPU_INT64, DFGPU_INT64};
  const dfgpu_expr_node c_pred[3] = {node(DFGPU_EXPR_COLUMN, 1, 0, 0), node(DFGPU_EXPR_LITERAL, 0, DFGPU_INT64, 1), node(DFGPU_EXPR_BINARY, DFGPU_OP_EQ, 0, 0)};
  dfgpu_pipeline* p1 = NULL;
  CHECK(dfgpu_pipeline_create(ctx, c_types, 2, c_pred, 3, NULL, 0, &p1));
  CHECK(dfgpu_pipeline_sink_build(p1, l1, 0, NULL, 0));
  dfgpu_column c_cols[2] = {col(DFGPU_INT64, c_custkey, NC), col(DFGPU_INT64, c_mktsegment, NC)};
  CHECK(dfgpu_pipeline_push_host(p1, c_cols, 2));
  CHECK(dfgpu_pipeline_finish(p1));
  dfgpu_pipeline_destroy(p1);

  /* ---- P2: orders -> join table with two payload fields and three accumulator words (row counter + a 128-bit sum) ---- */
  dfgpu_lookup_default_options(&lo);
  lo.n_acc_words = 3; lo.membership_filter = 1;                    /* the filter is what the lineitem scan tests first (dynamic filter pushdown) */
  const int32_t pay_types[2] = {DFGPU_DATE32, DFGPU_INT32};
  dfgpu_lookup* l2 = NULL;
  CHECK(dfgpu_lookup_create(ctx, DFGPU_INT64, pay_types, 2, &lo, &l2));
  const int32_t o_types[4] = {DFGPU_INT64, DFGPU_INT64, DFGPU_DATE32, DFGPU_INT32};
  const dfgpu_expr_node o_pred[3] = {node(DFGPU_EXPR_COLUMN, 2, 0, 0), node(DFGPU_EXPR_LITERAL, 0, DFGPU_DATE32, CUT), node(DFGPU_EXPR_BINARY, DFGPU_OP_LT, 0, 0)};
  dfgpu_pipeline_stage o_stage;
  o_stage.kind = DFGPU_STAGE_SEMI; o_stage.key_col = 1; o_stage.lookup = l1;
  dfgpu_pipeline* p2 = NULL;
  CHECK(dfgpu_pipeline_create(ctx, o_types, 4, o_pred, 3, &o_stage, 1, &p2));
  const int32_t o_payload[2] = {2, 3};
  CHECK(dfgpu_pipeline_sink_build(p2, l2, 0, o_payload, 2));
  dfgpu_column o_cols[4] = {col(DFGPU_INT64, o_orderkey, NO), col(DFGPU_INT64, o_custkey, NO), col(DFGPU_DATE32, o_orderdate, NO), col(DFGPU_INT32, o_shippriority, NO)};
  CHECK(dfgpu_pipeline_push_host(p2, o_cols, 4));
  CHECK(dfgpu_pipeline_finish(p2));
  dfgpu_pipeline_destroy(p2);

  /* ---- P3: lineitem -> filter -> probe -> SUM(l_extendedprice * (1 - l_discount)) into the matched record ---- */
  const int32_t l_types[4] = {DFGPU_INT64, money, money, DFGPU_DATE32};
  const dfgpu_expr_node l_pred[3] = {node(DFGPU_EXPR_COLUMN, 3, 0, 0), node(DFGPU_EXPR_LITERAL, 0, DFGPU_DATE32, CUT), node(DFGPU_EXPR_BINARY, DFGPU_OP_GT, 0, 0)};
  /* the planner coerces the Int64 literal 1 to Decimal128(20,0) (type_coercion/binary.rs:1265): (20,0) - (15,2) -> (23,2); (15,2) * (23,2) -> (38,4) */
  const dfgpu_expr_node revenue[5] = {node(DFGPU_EXPR_COLUMN, 1, 0, 0), node(DFGPU_EXPR_LITERAL, 0, DFGPU_DECIMAL128_TYPE(20, 0), 1), node(DFGPU_EXPR_COLUMN, 2, 0, 0),
                                      node(DFGPU_EXPR_BINARY, DFGPU_OP_MINUS, 0, 0), node(DFGPU_EXPR_BINARY, DFGPU_OP_MULTIPLY, 0, 0)};
  dfgpu_pipeline_stage l_stage;
  l_stage.kind = DFGPU_STAGE_INNER; l_stage.key_col = 0; l_stage.lookup = l2;
  dfgpu_pipeline* p3 = NULL;
  CHECK(dfgpu_pipeline_create(ctx, l_types, 4, l_pred, 3, &l_stage, 1, &p3));
  dfgpu_pipeline_agg agg;
  agg.func = DFGPU_AGG_SUM; agg.n_nodes = 5; agg.expr = revenue;
  const int32_t group_cols[3] = {0, 4, 5};                           /* l_orderkey + the two payload fields of the stage (virtual columns 4, 5) */
  CHECK(dfgpu_pipeline_sink_aggregate(p3, group_cols, 3, &agg, 1, DFGPU_AGG_SINGLE, 8192));
  dfgpu_column l_cols[4] = {col(DFGPU_INT64, l_orderkey, NL), col(money, l_extendedprice, NL), col(money, l_discount, NL), col(DFGPU_DATE32, l_shipdate, NL)};
  CHECK(dfgpu_pipeline_push_host(p3, l_cols, 4));
  CHECK(dfgpu_pipeline_finish(p3));

  /* ---- the same by nested loops ---- */
  int64_t exp_key[NO], exp_rev[NO];
  int32_t exp_date[NO], exp_prio[NO];
  int n_exp = 0;
  for (int o = 0; o < NO; ++o) {
    int cust_ok = 0;
    for (int c = 0; c < NC; ++c) if (c_custkey[c] == o_custkey[o] && c_mktsegment[c] == 1) cust_ok = 1;
    if (!cust_ok || !(o_orderdate[o] < CUT)) continue;
    int64_t rev = 0;
    int any = 0;
    for (int l = 0; l < NL; ++l)
      if (l_orderkey[l] == o_orderkey[o] && l_shipdate[l] > CUT) { rev += price_cents[l] * (100 - disc_pct[l]); any = 1; }   /* unscaled at scale 4 */
    if (any) { exp_key[n_exp] = o_orderkey[o]; exp_date[n_exp] = o_orderdate[o]; exp_prio[n_exp] = o_shippriority[o]; exp_rev[n_exp] = rev; n_exp++; }
  }

  /* ---- drain and compare (group order is unspecified, as after any hash aggregation) ---- */
  int matched = 0, rows_total = 0, ok = 1;
  for (;;) {
    dfgpu_batch* b = NULL;
    int rc = dfgpu_pipeline_next(p3, /*host=*/1, &b);
    if (rc == DFGPU_END) break;
    CHECK(rc);
    const int64_t rows = dfgpu_batch_num_rows(b);
    dfgpu_column out[4];
    for (int c = 0; c < 4; ++c) CHECK(dfgpu_batch_column(b, c, &out[c]));
    if (out[3].type != DFGPU_DECIMAL128_TYPE(38, 4)) { fprintf(stderr, "SUM type %d is not Decimal128(38,4)\n", out[3].type); ok = 0; }
    for (int64_t r = 0; r < rows; ++r, ++rows_total) {
      const int64_t key = ((const int64_t*)out[0].values)[r];
      const int32_t date = ((const int32_t*)out[1].values)[r], prio = ((const int32_t*)out[2].values)[r];
      const uint64_t lo_w = ((const uint64_t*)out[3].values)[2 * r], hi_w = ((const uint64_t*)out[3].values)[2 * r + 1];
      printf("row: %lld %d %d %llu.%04llu\n", (long long)key, date, prio, (unsigned long long)(lo_w / 10000), (unsigned long long)(lo_w % 10000));
      int found = 0;
      for (int e = 0; e < n_exp; ++e)
        if (exp_key[e] == key && exp_date[e] == date && exp_prio[e] == prio && (uint64_t)exp_rev[e] == lo_w && hi_w == 0) found = 1;
      matched += found;
    }
    dfgpu_batch_release(b);
  }
  if (rows_total != n_exp || matched != n_exp) ok = 0;
  printf("%d groups, %d expected, %s\n", rows_total, n_exp, ok ? "equal" : "MISMATCH");
  dfgpu_pipeline_destroy(p3);
  dfgpu_lookup_destroy(l2);
  dfgpu_lookup_destroy(l1);
  dfgpu_ctx_destroy(ctx);
  return ok ? 0 : 1;
}
