// select.cu — the selection kernels that sit beside filter in arrow-select:
//
//   nullif   arrow-select/src/nullif.rs:44-113   validity' = validity & !(right.values & right.validity); values shared
//   zip      arrow-select/src/zip.rs:99-226      out[i] = mask[i] is Some(true) ? truthy[i] : falsy[i]   (arrays and scalars)
//            ScalarZipper / PrimitiveScalarImpl  zip.rs:248-440 (both sides scalar)
//
// Both are HBM-bound streaming passes. nullif touches bitmaps only (3 x N/8 bytes). zip reads the mask word once per
// 64 rows (L1 broadcast), both value streams as 128-bit vectors and writes one; the output validity is a pure word-wise
// bitmap expression, computed by the thread that owns the first chunk of each 64-row word.
#include "bitmap.cuh"

namespace {

// ---- nullif ---------------------------------------------------------------------------------
struct NullifArgs {
  const uint8_t *lv;  // left validity or NULL
  int64_t loff;
  const uint8_t *rv;  // right (boolean) values
  int64_t roff;
  const uint8_t *rn;  // right validity or NULL
  int64_t rnoff;
  int64_t len;
  uint64_t *out;
  unsigned long long *res;
};

__global__ void __launch_bounds__(256) k_nullif(const NullifArgs a) {
  const int64_t words = (a.len + 63) >> 6;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned cnt = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    const int64_t pos = w << 6;
    const int64_t left = a.len - pos;
    const uint64_t lenmask = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
    uint64_t r = ld_bits64(a.rv, a.roff + pos, a.roff + a.len);
    if (a.rn) r &= ld_bits64(a.rn, a.rnoff + pos, a.rnoff + a.len);          // right.values() & nulls (nullif.rs:70-73)
    const uint64_t l = a.lv ? ld_bits64(a.lv, a.loff + pos, a.loff + a.len) : lenmask;
    const uint64_t v = l & ~r & lenmask;                                      // nullif.rs:78-103
    a.out[w] = v;
    cnt += __popcll(v);
  }
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)cnt);
}

// ---- zip ------------------------------------------------------------------------------------
struct ZipArgs {
  const uint8_t *mv;  // mask values
  int64_t moff;
  const uint8_t *mn;  // mask validity (only when it has nulls) or NULL
  int64_t mnoff;
  const uint8_t *t, *f;        // value buffers (array sides) or NULL (scalar side: see ts / fs)
  uint4 ts, fs;                // scalar value replicated over 16 bytes
  const uint8_t *tn, *fn;      // validity of the array sides or NULL
  int64_t tnoff, fnoff;
  int t_all, f_all;            // validity word of a side without a bitmap: all ones (1) / all zeros (0) — null scalars
  int64_t len;
  uint8_t *out;
  uint64_t *out_valid;         // NULL: no validity produced
  unsigned long long *res;
};

// mask of the bytes of a 32-bit word selected by `bits` (bit e -> element e of the chunk), W = element width
template <int W>
__device__ __forceinline__ uint32_t sub_mask(uint32_t bits, int s) {
  if constexpr (W == 1) {
    const uint32_t m4 = (bits >> (4 * s)) & 0xfu;
    return ((m4 * 0x00204081u) & 0x01010101u) * 0xffu;
  } else if constexpr (W == 2) {
    const uint32_t m2 = (bits >> (2 * s)) & 0x3u;
    return ((m2 & 1u) ? 0x0000ffffu : 0u) | ((m2 & 2u) ? 0xffff0000u : 0u);
  } else if constexpr (W == 4) {
    return ((bits >> s) & 1u) ? 0xffffffffu : 0u;
  } else if constexpr (W == 8) {
    return ((bits >> (s >> 1)) & 1u) ? 0xffffffffu : 0u;
  } else {
    return (bits & 1u) ? 0xffffffffu : 0u;
  }
}

// One thread per 16-byte chunk of the output. ROWS_PER_CHUNK = 16 / W (W <= 16); W = 32: two chunks per row.
template <int W>
__global__ void __launch_bounds__(256) k_zip(const ZipArgs a) {
  constexpr int RPC = W <= 16 ? 16 / W : 1;
  constexpr int CPR = W <= 16 ? 1 : W / 16;
  const int64_t chunks = (a.len * W + 15) / 16;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned cnt = 0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += stride) {
    const int64_t row0 = c * RPC / CPR;
    const int64_t wpos = row0 & ~(int64_t)63;
    uint64_t sel = ld_bits64(a.mv, a.moff + wpos, a.moff + a.len);
    if (a.mn) sel &= ld_bits64(a.mn, a.mnoff + wpos, a.mnoff + a.len);  // maybe_prep_null_mask_filter (zip.rs:662-671)
    const uint32_t bits = (uint32_t)(sel >> (row0 & 63)) & ((1u << RPC) - 1u);
    uint4 tv = a.ts, fv = a.fs;
    const bool last_partial = (c + 1) * 16 > a.len * W;  // the buffers are only guaranteed to hold len * W bytes
    if (!last_partial) {
      if (a.t) tv = ld_stream16(a.t + c * 16);
      if (a.f) fv = ld_stream16(a.f + c * 16);
    } else {
      const int nb = (int)(a.len * W - c * 16);
      if (a.t) { uint8_t *p = reinterpret_cast<uint8_t *>(&tv); for (int k = 0; k < nb; ++k) p[k] = __ldg(a.t + c * 16 + k); }
      if (a.f) { uint8_t *p = reinterpret_cast<uint8_t *>(&fv); for (int k = 0; k < nb; ++k) p[k] = __ldg(a.f + c * 16 + k); }
    }
    uint4 o;
    { const uint32_t m = sub_mask<W>(bits, 0); o.x = (tv.x & m) | (fv.x & ~m); }
    { const uint32_t m = sub_mask<W>(bits, 1); o.y = (tv.y & m) | (fv.y & ~m); }
    { const uint32_t m = sub_mask<W>(bits, 2); o.z = (tv.z & m) | (fv.z & ~m); }
    { const uint32_t m = sub_mask<W>(bits, 3); o.w = (tv.w & m) | (fv.w & ~m); }
    if (!last_partial) {
      st_stream16(a.out + c * 16, o);
    } else {
      const int nb = (int)(a.len * W - c * 16);
      const uint8_t *p = reinterpret_cast<const uint8_t *>(&o);
      for (int k = 0; k < nb; ++k) a.out[c * 16 + k] = p[k];
    }
    // the chunk that starts a 64-row word owns that word of the output validity
    if (a.out_valid && (row0 & 63) == 0 && (CPR == 1 || (c % CPR) == 0)) {
      const int64_t left = a.len - wpos;
      const uint64_t lenmask = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
      const uint64_t tw = a.tn ? ld_bits64(a.tn, a.tnoff + wpos, a.tnoff + a.len) : (a.t_all ? ~0ull : 0ull);
      const uint64_t fw = a.fn ? ld_bits64(a.fn, a.fnoff + wpos, a.fnoff + a.len) : (a.f_all ? ~0ull : 0ull);
      const uint64_t v = ((sel & tw) | (~sel & fw)) & lenmask;
      a.out_valid[wpos >> 6] = v;
      cnt += __popcll(v);
    }
  }
  if (a.out_valid) {
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)cnt);
  }
}

// element-wise variant for value pointers that are not 16-byte aligned (odd slices)
template <int W>
__global__ void __launch_bounds__(256) k_zip_elem(const ZipArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += stride) {
    bool s = ld_bit(a.mv, a.moff + i);
    if (a.mn) s = s && ld_bit(a.mn, a.mnoff + i);
    const uint8_t *src = s ? (a.t ? a.t + i * W : reinterpret_cast<const uint8_t *>(&a.ts)) : (a.f ? a.f + i * W : reinterpret_cast<const uint8_t *>(&a.fs));
    if (W == 32 && !(s ? a.t : a.f)) {  // scalar wider than the 16-byte register image: both halves hold the pattern halves
      // (32-byte scalars are passed through memory instead: see acu_zip)
    }
#pragma unroll
    for (int k = 0; k < W; ++k) a.out[i * W + k] = src[k];
    if (a.out_valid && (i & 63) == 0) {
      const int64_t left = a.len - i;
      const uint64_t lenmask = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
      uint64_t sel = ld_bits64(a.mv, a.moff + i, a.moff + a.len);
      if (a.mn) sel &= ld_bits64(a.mn, a.mnoff + i, a.mnoff + a.len);
      const uint64_t tw = a.tn ? ld_bits64(a.tn, a.tnoff + i, a.tnoff + a.len) : (a.t_all ? ~0ull : 0ull);
      const uint64_t fw = a.fn ? ld_bits64(a.fn, a.fnoff + i, a.fnoff + a.len) : (a.f_all ? ~0ull : 0ull);
      const uint64_t v = ((sel & tw) | (~sel & fw)) & lenmask;
      a.out_valid[i >> 6] = v;
      cnt += __popcll(v);
    }
  }
  if (a.out_valid) {
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)cnt);
  }
}

template <int W>
acu_status launch_zip(acu_ctx *ctx, const ZipArgs &a, bool vec) {
  if (vec) {
    const int64_t chunks = (a.len * W + 15) / 16;
    ACU_LAUNCH_TIMED(ctx, ACU_K_ARITH, (k_zip<W>), acu_grid(ctx, (chunks + 255) / 256, 32), 256, 0, a);
  } else {
    ACU_LAUNCH_TIMED(ctx, ACU_K_ARITH, (k_zip_elem<W>), acu_grid(ctx, (a.len + 255) / 256, 32), 256, 0, a);
  }
  return ACU_OK;
}

}  // namespace

extern "C" acu_status acu_nullif(acu_ctx *ctx, const acu_array *left, const acu_array *right, acu_array_out *out) {
  ACU_ENTER(ctx);
  if (left->len != right->len)  // nullif.rs:47-51
    return acu_fail(ctx, ACU_ERR_COMPUTE, -1, 0, 0, 0, "Cannot perform comparison operation on arrays of different length");
  const int64_t n = left->len;
  out->len = n;
  out->null_count = 0;
  out->has_validity = 0;
  if (n == 0) return ACU_OK;  // nullif.rs:54-56: the array is returned as it is
  NullifArgs a{};
  a.lv = left->validity;
  a.loff = left->validity_offset;
  a.rv = static_cast<const uint8_t *>(right->values);
  a.roff = right->values_offset;
  a.rn = right->validity;
  a.rnoff = right->validity_offset;
  a.len = n;
  a.out = reinterpret_cast<uint64_t *>(out->validity);
  a.res = acu_dres(ctx, 0);
  ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, k_nullif, acu_grid(ctx, ((n + 63) / 64 + 255) / 256, 16), 256, 0, a);
  ACU_TRY(acu_res_fetch(ctx));
  out->null_count = n - (int64_t)ctx->h_res[RES_COUNT];
  // data.nulls(Some(nulls)).build_unchecked(): the builder drops a NullBuffer without nulls (arrow-data/src/data.rs:2238-2251)
  out->has_validity = out->null_count > 0 ? 1 : 0;
  return ACU_OK;
}

extern "C" acu_status acu_zip(acu_ctx *ctx, int32_t elem_bytes, const acu_array *mask, const acu_array *truthy,
                              const acu_array *falsy, acu_array_out *out) {
  ACU_ENTER(ctx);
  const bool ts = truthy->is_scalar != 0, fs = falsy->is_scalar != 0;
  // zip.rs:123-145 (same texts for the ScalarZipper path, zip.rs:288-297)
  if (ts && truthy->len != 1) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "scalar arrays must have 1 element");
  if (!ts && truthy->len != mask->len) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "all arrays should have the same length");
  if (fs && falsy->len != 1) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "scalar arrays must have 1 element");
  if (!fs && falsy->len != mask->len) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "all arrays should have the same length");
  if (!(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 16 || elem_bytes == 32))
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "zip: unsupported element width %d", elem_bytes);
  const int64_t n = mask->len;
  out->len = n;
  out->null_count = 0;
  out->has_validity = 0;
  if (n == 0) return ACU_OK;
  acu_status st;
  const int64_t mnc = acu_resolve_null_count(ctx, mask, &st);
  ACU_TRY(st);
  const int64_t tnc = acu_resolve_null_count(ctx, truthy, &st);
  ACU_TRY(st);
  const int64_t fnc = acu_resolve_null_count(ctx, falsy, &st);
  ACU_TRY(st);
  ZipArgs a{};
  a.mv = static_cast<const uint8_t *>(mask->values);
  a.moff = mask->values_offset;
  if (mask->validity && mnc > 0) { a.mn = mask->validity; a.mnoff = mask->validity_offset; }
  a.len = n;
  a.out = static_cast<uint8_t *>(out->values);
  a.res = acu_dres(ctx, 0);
  a.t_all = a.f_all = 1;
  const uint8_t *tsrc = static_cast<const uint8_t *>(truthy->values), *fsrc = static_cast<const uint8_t *>(falsy->values);
  bool always_nulls = false;  // the result keeps its NullBuffer even without nulls
  bool produce_nulls;
  // scalar operands: fetch the single value (W bytes) and replicate it over a 16-byte register image
  uint8_t tval[32] = {0}, fval[32] = {0};
  if (ts) ACU_CUDA(ctx, cudaMemcpyAsync(tval, tsrc, (size_t)elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (fs) ACU_CUDA(ctx, cudaMemcpyAsync(fval, fsrc, (size_t)elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (ts || fs) ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const bool t_null = tnc > 0 && ts, f_null = fnc > 0 && fs;
  if (ts && fs) {
    // PrimitiveScalarImpl::create_output (zip.rs:392-440)
    if (!t_null && !f_null) {
      produce_nulls = false;
    } else if (!t_null && f_null) {  // every slot holds the truthy value, nulls = predicate
      memcpy(fval, tval, 32);
      a.t_all = 1; a.f_all = 0;
      produce_nulls = always_nulls = true;
    } else if (t_null && !f_null) {  // every slot holds the falsy value, nulls = !predicate
      memcpy(tval, fval, 32);
      a.t_all = 0; a.f_all = 1;
      produce_nulls = always_nulls = true;
    } else {                         // zeros, all null
      memset(tval, 0, 32);
      memset(fval, 0, 32);
      a.t_all = a.f_all = 0;
      produce_nulls = always_nulls = true;
    }
  } else {
    // zip_impl over MutableArrayData (zip.rs:156-226): values are copied blindly from the chosen side; a validity buffer is
    // built iff some input has nulls (arrow-data/src/transform/mod.rs:470) and dropped again when the result has none (:927-936)
    produce_nulls = tnc > 0 || fnc > 0;
    if (t_null) a.t_all = 0;
    else if (!ts && truthy->validity && tnc > 0) { a.tn = truthy->validity; a.tnoff = truthy->validity_offset; }
    if (f_null) a.f_all = 0;
    else if (!fs && falsy->validity && fnc > 0) { a.fn = falsy->validity; a.fnoff = falsy->validity_offset; }
  }
  if (!ts) a.t = tsrc;
  if (!fs) a.f = fsrc;
  auto replicate = [&](const uint8_t *v, uint4 *dst, int half) {
    uint8_t img[16];
    for (int k = 0; k < 16; ++k) img[k] = v[(elem_bytes > 16 ? half * 16 : 0) + (k % (elem_bytes > 16 ? 16 : elem_bytes))];
    memcpy(dst, img, 16);
  };
  if (produce_nulls) a.out_valid = reinterpret_cast<uint64_t *>(out->validity);
  // 32-byte scalars do not fit the register image: stage them in scratch as a 2-chunk pattern and treat them as arrays of period 1
  const bool wide_scalar = elem_bytes == 32 && (ts || fs);
  if (wide_scalar)
    return acu_fail(ctx, ACU_ERR_NOT_YET_IMPLEMENTED, -1, 0, 0, 0, "zip: 32-byte scalar operands are not supported");
  replicate(tval, &a.ts, 0);
  replicate(fval, &a.fs, 0);
  const bool vec = (a.t == nullptr || (uintptr_t)a.t % 16 == 0) && (a.f == nullptr || (uintptr_t)a.f % 16 == 0) && (uintptr_t)a.out % 16 == 0;
  ACU_TRY(acu_res_reset(ctx));
  switch (elem_bytes) {
    case 1: ACU_TRY(launch_zip<1>(ctx, a, vec)); break;
    case 2: ACU_TRY(launch_zip<2>(ctx, a, vec)); break;
    case 4: ACU_TRY(launch_zip<4>(ctx, a, vec)); break;
    case 8: ACU_TRY(launch_zip<8>(ctx, a, vec)); break;
    case 16: ACU_TRY(launch_zip<16>(ctx, a, vec)); break;
    default: ACU_TRY(launch_zip<32>(ctx, a, vec)); break;
  }
  ACU_TRY(acu_res_fetch(ctx));
  if (produce_nulls) {
    out->null_count = n - (int64_t)ctx->h_res[RES_COUNT];
    out->has_validity = (always_nulls || out->null_count > 0) ? 1 : 0;
  }
  return ACU_OK;
}
