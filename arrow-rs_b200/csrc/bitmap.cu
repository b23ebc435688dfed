// bitmap.cu — popcount / copy / AND of LSB-first bitmaps with arbitrary bit offsets.
// Reference: BooleanBuffer::count_set_bits, `&` (arrow-buffer/src/buffer/boolean.rs:702-719,
// arrow-buffer/src/buffer/ops.rs:149-170), BooleanArray::true_count
// (arrow-array/src/array/boolean_array.rs:175-187), NullBuffer::new (null.rs:41-44).
#include "bitmap.cuh"

// out[w] = a[w] (& b[w]) normalised to bit offset 0; optional popcount into res[RES_COUNT].
// One u64 word per thread, grid-stride; HBM-bound on len/8 bytes per operand.
__global__ void __launch_bounds__(256) k_bitmap_and(const uint8_t *__restrict__ a, int64_t aoff,
                                                    const uint8_t *__restrict__ b, int64_t boff,
                                                    int64_t len, uint64_t *__restrict__ out,
                                                    unsigned long long *__restrict__ res) {
  int64_t words = (len + 63) >> 6;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned cnt = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    uint64_t x = ld_bits64(a, aoff + (w << 6), aoff + len);
    if (b) x &= ld_bits64(b, boff + (w << 6), boff + len);
    if (out) out[w] = x;
    cnt += __popcll(x);
  }
  if (res) {
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(res + RES_COUNT, (unsigned long long)cnt);
  }
}

acu_status acu_bitmap_and_launch(acu_ctx *ctx, const uint8_t *a, int64_t aoff, const uint8_t *b,
                                 int64_t boff, int64_t len, uint64_t *out, bool count, unsigned long long *res) {
  if (len <= 0) return ACU_OK;
  int64_t words = (len + 63) >> 6;
  ACU_LAUNCH(ctx, k_bitmap_and, acu_grid(ctx, (words + 255) / 256, 8), 256, 0, a, aoff, b, boff, len, out,
             count ? (res ? res : ctx->d_res) : nullptr);
  return ACU_OK;
}

int64_t acu_resolve_null_count(acu_ctx *ctx, const acu_array *a, acu_status *st) {
  *st = ACU_OK;
  if (!a->validity) return 0;
  if (a->null_count >= 0) return a->null_count;
  int64_t len = a->is_scalar ? 1 : a->len;
  if (len == 0) return 0;
  if (ctx->async_on) {
    *st = acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0,
                   "null_count = -1 needs a device count (a synchronisation): pass the cached null_count inside an async section");
    return 0;
  }
  if ((*st = acu_res_reset(ctx)) != ACU_OK) return 0;
  if ((*st = acu_bitmap_and_launch(ctx, a->validity, a->validity_offset, nullptr, 0, len, nullptr, true)) != ACU_OK) return 0;
  if ((*st = acu_res_fetch(ctx)) != ACU_OK) return 0;
  return len - (int64_t)ctx->h_res[RES_COUNT];
}

extern "C" acu_status acu_bitmap_count(acu_ctx *ctx, const uint8_t *bits, int64_t offset,
                                       const uint8_t *validity, int64_t validity_offset, int64_t len,
                                       int64_t *out_count) {
  ACU_ENTER(ctx);
  *out_count = 0;
  if (len <= 0) return ACU_OK;
  ACU_TRY(acu_res_reset(ctx));
  ACU_TRY(acu_bitmap_and_launch(ctx, bits, offset, validity, validity_offset, len, nullptr, true));
  ACU_TRY(acu_res_fetch(ctx));
  *out_count = (int64_t)ctx->h_res[RES_COUNT];
  return ACU_OK;
}
