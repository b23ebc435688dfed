// views.cu — Utf8View / BinaryView buffer management for BatchCoalescer (arrow-select/src/coalesce/byte_view.rs):
// the device halves of InProgressByteViewArray. The policy (when to garbage-collect a source's data buffers, how large the
// next output buffer is, when the current one is full) stays on the host (host/arrow_cuda.hpp, acu/coalesce.py) exactly as
// in the reference; these entry points do the per-view work:
//   acu_view_bytes_used    GenericByteViewArray::total_buffer_bytes_used        arrow-array/src/array/byte_view_array.rs:749-761
//   acu_view_fit           the "copy as many views as fit" loop                 coalesce/byte_view.rs:259-271
//   acu_view_copy_strings  append_views_and_copy_strings_inner                  coalesce/byte_view.rs:298-354
//   acu_view_rebase        append_views_and_update_buffer_index                 coalesce/byte_view.rs:176-216
// A view is 16 bytes: length u32, then 12 inline bytes, or a 4-byte prefix + buffer index u32 + offset u32
// (arrow-data/src/byte_view.rs); MAX_INLINE_VIEW_LEN = 12.
#include "common.cuh"
#include "internal.cuh"

namespace {

constexpr uint32_t INLINE_MAX = 12;

// lens[i] = length of view i when it is stored out of line, else 0 (the bytes a compacted buffer needs for it)
__global__ void __launch_bounds__(256) k_view_long_lens(const uint4 *__restrict__ views, int64_t n, int64_t *__restrict__ lens) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t len = __ldg(reinterpret_cast<const uint32_t *>(views + i));
    lens[i] = len > INLINE_MAX ? (int64_t)len : 0;
  }
}

__global__ void __launch_bounds__(256) k_view_bytes_used(const uint4 *__restrict__ views, int64_t n, unsigned long long *__restrict__ res) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long sum = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t len = __ldg(reinterpret_cast<const uint32_t *>(views + i));
    if (len > INLINE_MAX) sum += len;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(ACU_FULL_MASK, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(res + RES_AUX0, sum);
}

// first view i with (capacity - bytes of the long views before i) < len(i): the reference's loop compares EVERY view's
// length (inline ones too) against what is left, and only long views consume capacity (coalesce/byte_view.rs:259-271)
__global__ void __launch_bounds__(256) k_view_fit(const uint4 *__restrict__ views, int64_t n, const int64_t *__restrict__ incl,
                                                  long long capacity, unsigned long long *__restrict__ res) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long first = ~0ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t len = __ldg(reinterpret_cast<const uint32_t *>(views + i));
    const long long before = incl[i] - (len > INLINE_MAX ? (long long)len : 0);
    if (capacity - before < (long long)len && (unsigned long long)i < first) first = (unsigned long long)i;
  }
  if (first != ~0ull) atomicMin(res + RES_ERR_INDEX, first);
}

// one warp per 32 views: lane j rewrites view j, the warp copies the bytes of every long view of the group
__global__ void __launch_bounds__(256) k_view_copy(const uint4 *__restrict__ views, int64_t n, const int64_t *__restrict__ incl,
                                                   const uint8_t *const *__restrict__ buffers, int n_buffers, uint32_t new_index,
                                                   uint8_t *__restrict__ dst, long long dst_len, uint4 *__restrict__ out_views,
                                                   unsigned long long *__restrict__ res) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp * 32; g < n; g += nwarps * 32) {
    const int64_t i = g + lane;
    uint4 v = make_uint4(0, 0, 0, 0);
    long long pos = 0;
    const uint8_t *src = nullptr;
    if (i < n) {
      v = __ldg(views + i);
      if (v.x > INLINE_MAX) {
        pos = dst_len + incl[i] - (long long)v.x;
        if (v.z < (uint32_t)n_buffers) src = buffers[v.z] + v.w;
        else atomicMin(res + RES_ERR_INDEX, (unsigned long long)i);  // a view that points outside its array's buffers
        out_views[i] = make_uint4(v.x, v.y, new_index, (uint32_t)pos);
      } else {
        out_views[i] = v;
      }
    }
    const uint32_t longs = __ballot_sync(ACU_FULL_MASK, src != nullptr);
    for (uint32_t m = longs; m; m &= m - 1) {
      const int j = __ffs(m) - 1;
      const uint32_t len = __shfl_sync(ACU_FULL_MASK, v.x, j);
      const long long p = __shfl_sync(ACU_FULL_MASK, pos, j);
      const uint8_t *s = reinterpret_cast<const uint8_t *>(__shfl_sync(ACU_FULL_MASK, (unsigned long long)(uintptr_t)src, j));
      uint8_t *d = dst + p;
      if (len >= 256 && (((uintptr_t)s ^ (uintptr_t)d) & 15) == 0) {  // long value, same 16-byte phase: head bytes, 128-bit body, tail
        const uint32_t head = (uint32_t)((16 - ((uintptr_t)d & 15)) & 15);
        if ((uint32_t)lane < head) d[lane] = s[lane];
        const uint32_t body = (len - head) >> 4;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s + head);
        uint4 *d4 = reinterpret_cast<uint4 *>(d + head);
        for (uint32_t c = lane; c < body; c += 32) d4[c] = s4[c];
        for (uint32_t b = head + (body << 4) + lane; b < len; b += 32) d[b] = s[b];
      } else {
        for (uint32_t b = lane; b < len; b += 32) d[b] = s[b];
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_view_rebase(const uint4 *__restrict__ views, int64_t n, uint32_t delta, uint4 *__restrict__ out_views) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint4 v = __ldg(views + i);
    if (v.x > INLINE_MAX) v.z += delta;
    out_views[i] = v;
  }
}

size_t scan_scratch_bytes(int64_t n) { return ((size_t)n + (size_t)n / 4096 + (size_t)n / (4096 * 4096) + 64) * 8; }

}  // namespace

extern "C" {

acu_status acu_view_bytes_used(acu_ctx *ctx, const void *views, int64_t n, int64_t *out_total) {
  ACU_ENTER(ctx);
  *out_total = 0;
  if (n <= 0) return ACU_OK;
  ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH(ctx, k_view_bytes_used, acu_grid(ctx, (n + 255) / 256, 16), 256, 0, static_cast<const uint4 *>(views), n, ctx->d_res);
  ACU_TRY(acu_res_fetch(ctx));
  *out_total = (int64_t)ctx->h_res[RES_AUX0];
  return ACU_OK;
}

acu_status acu_view_fit(acu_ctx *ctx, const void *views, int64_t n, int64_t remaining_capacity, int64_t *out_views, int64_t *out_bytes) {
  ACU_ENTER(ctx);
  *out_views = 0;
  *out_bytes = 0;
  if (n <= 0) return ACU_OK;
  void *scratch;
  ACU_TRY(acu_scratch(ctx, scan_scratch_bytes(n), &scratch));
  int64_t *incl = static_cast<int64_t *>(scratch);
  ACU_TRY(acu_res_reset(ctx));
  const int grid = acu_grid(ctx, (n + 255) / 256, 16);
  ACU_LAUNCH(ctx, k_view_long_lens, grid, 256, 0, static_cast<const uint4 *>(views), n, incl);
  ACU_TRY(acu_scan_inclusive_i64(ctx, incl, n, incl + n));
  ACU_LAUNCH(ctx, k_view_fit, grid, 256, 0, static_cast<const uint4 *>(views), n, incl, (long long)remaining_capacity, ctx->d_res);
  ACU_TRY(acu_res_fetch(ctx));
  const int64_t first = ctx->h_res[RES_ERR_INDEX] == ~0ull ? n : (int64_t)ctx->h_res[RES_ERR_INDEX];
  *out_views = first;
  if (first > 0) {  // bytes of the long views among [0, first)
    int64_t b = 0;
    ACU_CUDA(ctx, cudaMemcpyAsync(&b, incl + (first - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out_bytes = b;
  }
  return ACU_OK;
}

acu_status acu_view_copy_strings(acu_ctx *ctx, const void *views, int64_t n, const uint8_t *const *buffers, int32_t n_buffers,
                                 uint32_t new_buffer_index, uint8_t *dst, int64_t dst_len, int64_t dst_capacity, void *out_views,
                                 int64_t *out_bytes) {
  ACU_ENTER(ctx);
  *out_bytes = 0;
  if (n <= 0) return ACU_OK;
  if (n_buffers < 0) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "view_copy_strings: n_buffers %d", n_buffers);
  void *scratch;
  ACU_TRY(acu_scratch(ctx, scan_scratch_bytes(n) + (size_t)(n_buffers + 1) * sizeof(void *) + 256, &scratch));
  int64_t *incl = static_cast<int64_t *>(scratch);
  const uint8_t **table = reinterpret_cast<const uint8_t **>(static_cast<uint8_t *>(scratch) + ((scan_scratch_bytes(n) + 255) & ~(size_t)255));
  if (n_buffers) ACU_CUDA(ctx, cudaMemcpyAsync(table, buffers, (size_t)n_buffers * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
  const int grid = acu_grid(ctx, (n + 255) / 256, 16);
  ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH(ctx, k_view_long_lens, grid, 256, 0, static_cast<const uint4 *>(views), n, incl);
  ACU_TRY(acu_scan_inclusive_i64(ctx, incl, n, incl + n));
  // the copy is only queued once the total is known to fit (a sizing bug of the caller must not write past dst)
  int64_t total = 0;
  ACU_CUDA(ctx, cudaMemcpyAsync(&total, incl + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
  ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (dst_len + total > dst_capacity)
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, (uint64_t)dst_capacity,
                    "view_copy_strings: %lld bytes do not fit a buffer of capacity %lld holding %lld", (long long)total, (long long)dst_capacity,
                    (long long)dst_len);
  if (dst_len + total > (int64_t)UINT32_MAX)
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "view_copy_strings: a view offset is 32 bits, buffer of %lld bytes", (long long)(dst_len + total));
  ACU_LAUNCH_TIMED(ctx, ACU_K_BYTES, k_view_copy, acu_grid(ctx, (n + 255) / 256, 16), 256, 0, static_cast<const uint4 *>(views), n, incl, table, (int)n_buffers,
                   new_buffer_index, dst, (long long)dst_len, static_cast<uint4 *>(out_views), ctx->d_res);
  ACU_TRY(acu_res_fetch(ctx));
  if (ctx->h_res[RES_ERR_INDEX] != ~0ull)
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, (int64_t)ctx->h_res[RES_ERR_INDEX], 0, 0, (uint64_t)n_buffers,
                    "view %lld refers to a data buffer the array does not have (%d buffers)", (long long)ctx->h_res[RES_ERR_INDEX], n_buffers);
  *out_bytes = total;
  return ACU_OK;
}

acu_status acu_view_rebase(acu_ctx *ctx, const void *views, int64_t n, uint32_t delta, void *out_views) {
  ACU_ENTER(ctx);
  if (n <= 0) return ACU_OK;
  ACU_LAUNCH(ctx, k_view_rebase, acu_grid(ctx, (n + 255) / 256, 16), 256, 0, static_cast<const uint4 *>(views), n, delta, static_cast<uint4 *>(out_views));
  ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ACU_OK;
}

}  // extern "C"
