// internal.cuh — cross-translation-unit internals (not part of the C ABI).
#pragma once
#include "common.cuh"

struct acu_filter_plan;
const uint64_t *acu_plan_mask(const acu_filter_plan *p);      // normalised mask words (padded to x32)
const uint64_t *acu_plan_tile_off(const acu_filter_plan *p);  // exclusive output offset per 1024-row tile
int64_t acu_plan_n_tiles(const acu_filter_plan *p);
int64_t acu_plan_n_words_padded(const acu_filter_plan *p);

// FilterPredicate::filter_nulls (filter.rs:512-533) for any array kind.
acu_status acu_filter_nulls_internal(acu_ctx *ctx, const acu_filter_plan *plan, const acu_array *a,
                                     acu_array_out *out);

// Shared front end of take_primitive / take_boolean / take_bytes (take.cu).
acu_status acu_take_common(acu_ctx *ctx, int32_t elem_bytes, const acu_array *values, bool boolean_values,
                           const acu_array *indices, acu_dtype index_dtype, int32_t check_bounds,
                           acu_array_out *out);
