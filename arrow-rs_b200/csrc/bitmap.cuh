// bitmap.cuh — internal bitmap helpers shared between translation units.
#pragma once
#include "common.cuh"

// out = a (& b) normalised to bit offset 0 (out may be NULL); if `count`, the popcount of
// the result is atomically added to res[RES_COUNT] (res == NULL: result block 0; the caller
// resets / fetches).
acu_status acu_bitmap_and_launch(acu_ctx *ctx, const uint8_t *a, int64_t aoff, const uint8_t *b,
                                 int64_t boff, int64_t len, uint64_t *out, bool count,
                                 unsigned long long *res = nullptr);
