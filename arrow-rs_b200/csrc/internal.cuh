// internal.cuh — cross-translation-unit internals (not part of the C ABI).
#pragma once
#include "common.cuh"

struct acu_filter_plan;
const uint64_t *acu_plan_mask(const acu_filter_plan *p);
const uint32_t *acu_plan_tile_local(const acu_filter_plan *p);
const uint32_t *acu_plan_tile_count(const acu_filter_plan *p);
const uint64_t *acu_plan_chunk_offset(const acu_filter_plan *p);
int64_t acu_plan_n_tiles(const acu_filter_plan *p);

// FilterPredicate::filter_nulls (filter.rs:512-533) for any array kind.
acu_status acu_filter_nulls_internal(acu_ctx *ctx, const acu_filter_plan *plan, const acu_array *a,
                                     acu_array_out *out);

// Shared front end of take_primitive / take_boolean / take_bytes (take.cu).
acu_status acu_take_common(acu_ctx *ctx, int32_t elem_bytes, const acu_array *values, bool boolean_values,
                           const acu_array *indices, acu_dtype index_dtype, int32_t check_bounds,
                           acu_array_out *out);
