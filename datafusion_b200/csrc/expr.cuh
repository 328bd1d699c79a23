// expr.cuh — PhysicalExpr::evaluate on device columns (defined in filter.cu), shared with the join filter.
#pragma once
#include "batch.cuh"

namespace dfgpu {

struct ExprPlan {
  std::vector<dfgpu_expr_node> nodes;
  std::vector<int> in_type, out_type;
  int root_type = 0;
};
// type inference + validation of a post-order program against a schema
ExprPlan plan_expr(const int32_t* schema_types, int n_cols, const dfgpu_expr_node* nodes, int n_nodes);

uint64_t literal_bits(const dfgpu_expr_node& nd);   // literal value as the interpreter's 64-bit stack payload

struct EvalResult { DCol column; DevBuf select_words; };
// want_column: materialise the value column; want_select: selection words (valid AND true) for predicates
EvalResult evaluate_expr(dfgpu_ctx* ctx, const ExprPlan& plan, const std::vector<DCol>& cols, int64_t n, bool want_column, bool want_select);

}  // namespace dfgpu
