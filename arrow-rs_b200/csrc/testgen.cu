// testgen.cu — deterministic synthetic inputs generated on the device for bench.py, tools/ and the tests.
// NOT part of the product: built into libarrow_cuda_testgen.so (include/arrow_cuda_testgen.h), which links against
// libarrow_cuda.so for the context / launch helpers. Host twin: oracle/oracle.cpp orc_generate_*.
#include "common.cuh"
#include "../../include/arrow_cuda_testgen.h"

__global__ void __launch_bounds__(256) k_generate_values(int kind, uint64_t seed, int64_t first_row,
                                                         uint64_t param, void *out, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t r = splitmix64(seed ^ (uint64_t)(first_row + i));
    switch (kind) {
      case 0: static_cast<uint64_t *>(out)[i] = r; break;
      case 1: static_cast<int64_t *>(out)[i] = (int64_t)(r >> 2) - ((int64_t)1 << 61); break;
      case 2: static_cast<double *>(out)[i] = __dadd_rn(__dmul_rn((double)(r >> 11), 2.0e6 / 9007199254740992.0), -1.0e6); break;
      case 3: static_cast<uint32_t *>(out)[i] = (uint32_t)__umul64hi(r, param); break;
      default: static_cast<int32_t *>(out)[i] = (int32_t)__umul64hi(r, param); break;
    }
  }
}

__global__ void __launch_bounds__(256) k_generate_bits(uint64_t seed, int64_t first_row, uint64_t thr,
                                                       int all, uint64_t *out, int64_t n) {
  // one thread per output byte-lane: warp ballot packs 32 rows at a time
  int64_t words = (n + 63) / 64;
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < words; w += nwarps) {
    int64_t i0 = w * 64 + lane, i1 = i0 + 32;
    bool b0 = i0 < n && (all || splitmix64(seed ^ (uint64_t)(first_row + i0)) < thr);
    bool b1 = i1 < n && (all || splitmix64(seed ^ (uint64_t)(first_row + i1)) < thr);
    uint32_t lo = __ballot_sync(ACU_FULL_MASK, b0), hi = __ballot_sync(ACU_FULL_MASK, b1);
    if (lane == 0) out[w] = (uint64_t)lo | ((uint64_t)hi << 32);
  }
}

extern "C" acu_status acu_generate_values(acu_ctx *ctx, int32_t kind, uint64_t seed, int64_t first_row,
                                          uint64_t param, void *out, int64_t n) {
  ACU_ENTER(ctx);
  if (kind < 0 || kind > 4) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "unknown generator kind %d", kind);
  if (n <= 0) return ACU_OK;
  ACU_LAUNCH(ctx, k_generate_values, acu_grid(ctx, (n + 255) / 256, 16), 256, 0, kind, seed, first_row, param, out, n);
  return ACU_OK;
}

extern "C" acu_status acu_generate_bits(acu_ctx *ctx, uint64_t seed, int64_t first_row, double p,
                                        uint8_t *out_bits, int64_t n) {
  ACU_ENTER(ctx);
  if (n <= 0) return ACU_OK;
  uint64_t thr = p >= 1.0 ? ~0ull : (uint64_t)(p * 18446744073709551616.0);
  int64_t words = (n + 63) / 64;
  ACU_LAUNCH(ctx, k_generate_bits, acu_grid(ctx, (words + 7) / 8, 16), 256, 0, seed, first_row, thr,
             p >= 1.0 ? 1 : 0, reinterpret_cast<uint64_t *>(out_bits), n);
  return ACU_OK;
}
