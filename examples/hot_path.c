/* examples/hot_path.c — the hot path from plain C through include/arrow_cuda.h:
 *
 *     filter(col, pred) -> take(col, selected rows) -> add(a, b) -> sum(taken)
 *
 * on device-generated synthetic data (the generators of SURVEY.md §8(d)). Build:
 *     gcc -std=c11 -Iinclude examples/hot_path.c -Larrow-rs_b200 -larrow_cuda -Wl,-rpath,$PWD/arrow-rs_b200 -o hot_path
 * Every call returns an acu_status; acu_last_error(ctx)->message holds the reference's ArrowError text. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "arrow_cuda.h"
#include "arrow_cuda_testgen.h" /* synthetic inputs: test support, not part of the product library */

#define CHECK(call)                                                                              \
  do {                                                                                           \
    acu_status st_ = (call);                                                                     \
    if (st_ != ACU_OK) {                                                                         \
      fprintf(stderr, "%s failed: %d %s\n", #call, (int)st_, ctx ? acu_last_error(ctx)->message : ""); \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 10 * 1000 * 1000;
  acu_ctx *ctx = NULL;
  CHECK(acu_ctx_create(0, &ctx)); /* no CPU fallback: fails without a GPU */
  const size_t bb = acu_bitmap_bytes(n);
  void *col, *a, *b, *col_valid, *a_valid, *b_valid, *pred_bits;
  CHECK(acu_malloc(ctx, (size_t)n * 8, &col));
  CHECK(acu_malloc(ctx, (size_t)n * 8, &a));
  CHECK(acu_malloc(ctx, (size_t)n * 8, &b));
  CHECK(acu_malloc(ctx, bb, &col_valid));
  CHECK(acu_malloc(ctx, bb, &a_valid));
  CHECK(acu_malloc(ctx, bb, &b_valid));
  CHECK(acu_malloc(ctx, bb, &pred_bits));
  CHECK(acu_generate_values(ctx, 0, 42, 0, 0, col, n)); /* Int64, full range */
  CHECK(acu_generate_values(ctx, 2, 42, 0, 0, a, n));   /* Float64 in [-1e6, 1e6) */
  CHECK(acu_generate_values(ctx, 2, 43, 0, 0, b, n));
  CHECK(acu_generate_bits(ctx, 44, 0, 0.95, (uint8_t *)col_valid, n)); /* 5 % nulls */
  CHECK(acu_generate_bits(ctx, 45, 0, 0.95, (uint8_t *)a_valid, n));
  CHECK(acu_generate_bits(ctx, 46, 0, 0.95, (uint8_t *)b_valid, n));
  CHECK(acu_generate_bits(ctx, 47, 0, 0.10, (uint8_t *)pred_bits, n)); /* 10 % selected */

  acu_array pred = {pred_bits, 0, NULL, 0, n, 0, 0, 0};
  acu_array values = {col, 0, (const uint8_t *)col_valid, 0, n, -1, 0, 0}; /* null_count unknown: counted on device */
  acu_array fa = {a, 0, (const uint8_t *)a_valid, 0, n, -1, 0, 0};
  acu_array fb = {b, 0, (const uint8_t *)b_valid, 0, n, -1, 0, 0};

  /* FilterBuilder::new(&pred).optimize().build() */
  acu_filter_plan *plan = NULL;
  CHECK(acu_filter_plan_create(ctx, &pred, &plan));
  const int64_t count = acu_filter_plan_count(plan);
  void *f_vals, *f_valid, *idx, *t_vals, *t_valid, *s_vals, *s_valid;
  CHECK(acu_malloc(ctx, (size_t)count * 8 + 8, &f_vals));
  CHECK(acu_malloc(ctx, acu_bitmap_bytes(count) + 8, &f_valid));
  CHECK(acu_malloc(ctx, (size_t)count * 4 + 8, &idx));
  CHECK(acu_malloc(ctx, (size_t)count * 8 + 8, &t_vals));
  CHECK(acu_malloc(ctx, acu_bitmap_bytes(count) + 8, &t_valid));
  CHECK(acu_malloc(ctx, (size_t)n * 8, &s_vals));
  CHECK(acu_malloc(ctx, bb, &s_valid));

  acu_array_out filtered = {f_vals, (uint8_t *)f_valid, 0, 0, 0, 0};
  CHECK(acu_filter_primitive(ctx, plan, 8, &values, &filtered));          /* arrow::compute::filter */
  CHECK(acu_filter_plan_indices(ctx, plan, ACU_U32, idx));                /* the selected rows, ascending */
  acu_filter_plan_destroy(ctx, plan);

  acu_array indices = {idx, 0, NULL, 0, count, 0, 0, 0};
  acu_array_out taken = {t_vals, (uint8_t *)t_valid, 0, 0, 0, 0};
  CHECK(acu_take_primitive(ctx, 8, &values, &indices, ACU_U32, 0, &taken)); /* arrow::compute::take */

  acu_array_out sum_ab = {s_vals, (uint8_t *)s_valid, 0, 0, 0, 0};
  CHECK(acu_arith(ctx, ACU_F64, ACU_ADD, &fa, &fb, &sum_ab));             /* kernels::numeric::add */

  acu_array t_in = {taken.values, 0, taken.has_validity ? taken.validity : NULL, 0, taken.len, taken.has_validity ? taken.null_count : 0, 0, 0};
  uint64_t bits = 0;
  int64_t valid_rows = 0;
  CHECK(acu_aggregate(ctx, ACU_I64, ACU_SUM, &t_in, &bits, &valid_rows)); /* arrow::compute::sum */

  /* filter and take of the selected rows must agree: same length, same null count */
  if (filtered.len != taken.len || filtered.null_count != taken.null_count) {
    fprintf(stderr, "filter/take disagree\n");
    return 1;
  }
  printf("rows %" PRId64 ", selected %" PRId64 " (%" PRId64 " null), add nulls %" PRId64 ", sum(taken) = %" PRId64 " over %" PRId64 " valid rows\n", n,
         filtered.len, filtered.null_count, sum_ab.null_count, (int64_t)bits, valid_rows);
  void *all[] = {col, a, b, col_valid, a_valid, b_valid, pred_bits, f_vals, f_valid, idx, t_vals, t_valid, s_vals, s_valid};
  for (size_t i = 0; i < sizeof all / sizeof all[0]; ++i) acu_free(ctx, all[i]);
  acu_ctx_destroy(ctx);
  return 0;
}
