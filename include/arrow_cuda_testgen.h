/* arrow_cuda_testgen.h — synthetic-input generators for the benchmarks and tests (libarrow_cuda_testgen.so).
 * Test / benchmark support only: nothing in the product library (libarrow_cuda.so, include/arrow_cuda.h) depends on it. */
#ifndef ARROW_CUDA_TESTGEN_H
#define ARROW_CUDA_TESTGEN_H

#include "arrow_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic synthetic inputs generated on the device:
 * element i = f(splitmix64(seed ^ (first_row + i))). The same generator exists on the
 * host in oracle/ so that any row range can be spot-checked.
 *   kind 0: raw 64-bit values (Int64 full range)
 *   kind 1: Int64 uniform in [-2^61, 2^61)
 *   kind 2: Float64 uniform in [-1e6, 1e6)
 *   kind 3: UInt32 uniform in [0, param)         (take indices, distribution B)
 *   kind 4: Int32 uniform in [0, param)
 * acu_generate_bits: bit i = splitmix64(seed ^ (first_row+i)) < p * 2^64. */
acu_status acu_generate_values(acu_ctx *ctx, int32_t kind, uint64_t seed, int64_t first_row,
                               uint64_t param, void *out, int64_t n);
acu_status acu_generate_bits(acu_ctx *ctx, uint64_t seed, int64_t first_row, double p,
                             uint8_t *out_bits, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
