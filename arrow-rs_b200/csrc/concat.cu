// concat.cu — the building blocks of coalescing: append a row range of a source column to an
// in-progress destination column at an arbitrary row position (SURVEY.md §8(f) rank 1:
// BatchCoalescer::push_batch / push_batch_with_filter, arrow-select/src/coalesce.rs:258-533, whose
// InProgressArray::copy_rows appends `len` rows starting at `offset` of the current source).
//
//   values of fixed width  : a device-to-device copy (acu_memcpy_d2d)
//   validity / boolean bits: acu_bitmap_copy — bits [src_off, src_off+len) to [dst_off, dst_off+len),
//                            every other destination bit preserved
//   all-valid ranges       : acu_bitmap_fill
//   Utf8 offsets           : acu_offsets_append — rebased on the destination's running byte total
//
// One thread per destination u64 word; boundary words are merged with atomicOr / atomicAnd so that
// two appends never need the destination to start on a word.
#include "bitmap.cuh"

namespace {

// dst bits [doff, doff+len) = src bits [soff, soff+len); popcount of the copied bits -> res[RES_COUNT]
__global__ void __launch_bounds__(256) k_bitmap_copy(const uint8_t *__restrict__ src, int64_t soff, unsigned long long *__restrict__ dst,
                                                     int64_t doff, int64_t len, unsigned long long *__restrict__ res) {
  const int64_t w0 = doff >> 6, w1 = (doff + len - 1) >> 6;  // destination words touched
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned cnt = 0;
  for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += stride) {
    const int64_t lo = w << 6;                                  // first destination bit of this word
    const int64_t b0 = lo < doff ? doff : lo;                   // range of this word that is written
    const int64_t b1 = lo + 64 > doff + len ? doff + len : lo + 64;
    const unsigned sh = (unsigned)(b0 - lo);
    const int n = (int)(b1 - b0);
    uint64_t bits = ld_bits64(src, soff + (b0 - doff), soff + len);  // n valid bits (zero beyond the source range)
    if (n < 64) bits &= (1ull << n) - 1ull;
    cnt += __popcll(bits);
    const uint64_t mask = (n == 64 ? ~0ull : ((1ull << n) - 1ull)) << sh;
    if (n == 64) {
      dst[w] = bits;
    } else {  // boundary word: clear the range, then set — other bits (earlier appends) are untouched
      atomicAnd(dst + w, ~mask);
      atomicOr(dst + w, bits << sh);
    }
  }
  if (res) {
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(res + RES_COUNT, (unsigned long long)cnt);
  }
}

__global__ void __launch_bounds__(256) k_bitmap_fill(unsigned long long *__restrict__ dst, int64_t doff, int64_t len, int value) {
  const int64_t w0 = doff >> 6, w1 = (doff + len - 1) >> 6;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += stride) {
    const int64_t lo = w << 6;
    const int64_t b0 = lo < doff ? doff : lo, b1 = lo + 64 > doff + len ? doff + len : lo + 64;
    const int n = (int)(b1 - b0);
    const uint64_t mask = (n == 64 ? ~0ull : ((1ull << n) - 1ull)) << (unsigned)(b0 - lo);
    if (n == 64) dst[w] = value ? ~0ull : 0ull;
    else if (value) atomicOr(dst + w, mask);
    else atomicAnd(dst + w, ~mask);
  }
}

// dst[dfirst + j] = base + src[first + j] - src[first], j = 0 .. count (count + 1 entries); the first entry
// past `limit` goes to res[RES_ERR_INDEX]; res[RES_AUX0] = src[first], res[RES_AUX1] = src[first + count]
template <class O>
__global__ void __launch_bounds__(256) k_offsets_append(const O *__restrict__ src, int64_t first, int64_t count, int64_t base,
                                                        O *__restrict__ dst, int64_t dfirst, int64_t limit, unsigned long long *__restrict__ res) {
  const int64_t s0 = (int64_t)__ldg(src + first);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long err = ~0ull;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= count; j += stride) {
    const int64_t v = base + ((int64_t)__ldg(src + first + j) - s0);
    if (v > limit && (unsigned long long)j < err) err = (unsigned long long)j;
    dst[dfirst + j] = (O)v;
  }
  if (err != ~0ull) atomicMin(res + RES_ERR_INDEX, err);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    res[RES_AUX0] = (unsigned long long)s0;
    res[RES_AUX1] = (unsigned long long)(int64_t)__ldg(src + first + count);
  }
}

}  // namespace

extern "C" acu_status acu_bitmap_copy(acu_ctx *ctx, const uint8_t *src, int64_t src_offset, uint8_t *dst, int64_t dst_offset, int64_t len,
                                      int64_t *out_set_bits) {
  ACU_ENTER(ctx);
  if (out_set_bits) *out_set_bits = 0;
  if (len <= 0) return ACU_OK;
  if (((uintptr_t)dst & 7) != 0) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "bitmap_copy: destination must be 8-byte aligned");
  const int64_t words = ((dst_offset + len - 1) >> 6) - (dst_offset >> 6) + 1;
  if (out_set_bits) ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH(ctx, k_bitmap_copy, acu_grid(ctx, (words + 255) / 256, 8), 256, 0, src, src_offset, reinterpret_cast<unsigned long long *>(dst),
             dst_offset, len, out_set_bits ? ctx->d_res : nullptr);
  if (out_set_bits) {
    ACU_TRY(acu_res_fetch(ctx));
    *out_set_bits = (int64_t)ctx->h_res[RES_COUNT];
  }
  return ACU_OK;
}

extern "C" acu_status acu_bitmap_fill(acu_ctx *ctx, uint8_t *dst, int64_t dst_offset, int64_t len, int32_t value) {
  ACU_ENTER(ctx);
  if (len <= 0) return ACU_OK;
  if (((uintptr_t)dst & 7) != 0) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "bitmap_fill: destination must be 8-byte aligned");
  const int64_t words = ((dst_offset + len - 1) >> 6) - (dst_offset >> 6) + 1;
  ACU_LAUNCH(ctx, k_bitmap_fill, acu_grid(ctx, (words + 255) / 256, 8), 256, 0, reinterpret_cast<unsigned long long *>(dst), dst_offset, len, (int)value);
  return ACU_OK;
}

extern "C" acu_status acu_offsets_append(acu_ctx *ctx, int32_t offset_bytes, const void *src_offsets, int64_t first, int64_t count,
                                         int64_t base, void *dst_offsets, int64_t dst_first, int64_t *out_src_begin, int64_t *out_src_end) {
  ACU_ENTER(ctx);
  if (offset_bytes != 4 && offset_bytes != 8) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "offset width must be 4 or 8");
  if (count < 0) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "offsets_append: negative count");
  ACU_TRY(acu_res_reset(ctx));
  const int grid = acu_grid(ctx, (count + 1 + 255) / 256, 8);
  if (offset_bytes == 4)
    ACU_LAUNCH(ctx, k_offsets_append<int32_t>, grid, 256, 0, static_cast<const int32_t *>(src_offsets), first, count, base,
               static_cast<int32_t *>(dst_offsets), dst_first, (int64_t)INT32_MAX, ctx->d_res);
  else
    ACU_LAUNCH(ctx, k_offsets_append<int64_t>, grid, 256, 0, static_cast<const int64_t *>(src_offsets), first, count, base,
               static_cast<int64_t *>(dst_offsets), dst_first, INT64_MAX, ctx->d_res);
  ACU_TRY(acu_res_fetch(ctx));
  const int64_t s0 = (int64_t)ctx->h_res[RES_AUX0], s1 = (int64_t)ctx->h_res[RES_AUX1];
  if (out_src_begin) *out_src_begin = s0;
  if (out_src_end) *out_src_end = s1;
  if (ctx->h_res[RES_ERR_INDEX] != ~0ull) {  // the appended values no longer fit the offset type
    const long long total = (long long)(base + (s1 - s0));
    return acu_fail(ctx, ACU_ERR_OFFSET_OVERFLOW, (int64_t)ctx->h_res[RES_ERR_INDEX], 0, 0, (uint64_t)total, "%lld", total);
  }
  return ACU_OK;
}
