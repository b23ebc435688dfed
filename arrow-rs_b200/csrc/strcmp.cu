// strcmp.cu — arrow-ord/src/cmp.rs on variable-width operands:
//
//   eq / neq / lt / lt_eq / gt / gt_eq / distinct / not_distinct over
//     GenericByteArray      (Utf8, Binary: i32 offsets; LargeUtf8, LargeBinary: i64)   ArrayOrd cmp.rs:783-801
//     GenericByteViewArray  (Utf8View, BinaryView)                                      ArrayOrd cmp.rs:803-898
//   and the short-constant fast path of views, eq_inline_scalar                          cmp.rs:282-300, :405-435
//
// compare_op's null handling (cmp.rs:319-381) is the same as for primitives (elementwise.cu): value bits are computed at
// every slot, validity = union of the inputs' (folded into the values for distinct / not_distinct), a null scalar makes
// the result all-null. One thread per row, 4 rows per lane in flight, result bits packed with a warp ballot (lane == bit),
// lane 0 of each 32-row group owns the group's 32-bit value and validity words.
// Roofline: 2 offsets + the compared bytes per side (byte arrays), 16 B per view (views); HBM-bound.
#include "bitmap.cuh"
#include "internal.cuh"

namespace {

enum { SFOLD_NONE = 0, SFOLD_DISTINCT = 1, SFOLD_NOT_DISTINCT = 2 };

struct RowCmpCommon {
  int64_t n;
  const uint8_t *av, *bv;  // validity of the two (possibly swapped) operands, NULL = no nulls
  int64_t aoff, boff;
  int a_scalar, b_scalar;
  int a_null_scalar, b_null_scalar;
  int lt;                  // 1: is_lt(a, b), 0: is_eq(a, b)
  int neg, fold;
  uint32_t *out_bits, *out_valid;
  unsigned long long *res;
};

struct BytesOperand {
  const void *offs;
  const uint8_t *data;
  int ob;
};

struct ViewOperand {
  const uint4 *views;
  const uint8_t *const *buffers;  // device array of device pointers
  int n_buffers;
};

__device__ __forceinline__ int64_t ld_offset(const void *offs, int ob, int64_t i) {
  return ob == 4 ? (int64_t)__ldg(static_cast<const int32_t *>(offs) + i) : __ldg(static_cast<const int64_t *>(offs) + i);
}

// Up to 8 bytes of p[0 .. nb) as a little-endian u64, zero above nb: aligned 8-byte loads that contain a requested byte.
__device__ __forceinline__ uint64_t ld_upto8(const uint8_t *__restrict__ p, uint32_t nb) {
  const uintptr_t addr = (uintptr_t)p;
  const uint64_t *q = reinterpret_cast<const uint64_t *>(addr & ~(uintptr_t)7);
  const uint32_t sh = (uint32_t)(addr & 7u) * 8u;
  uint64_t lo = 0, hi = 0;
  if (nb) lo = __ldg(q);
  if (sh + nb * 8u > 64u) hi = __ldg(q + 1);
  const uint64_t w = (lo >> sh) | ((hi << 1) << (63u - sh));
  return w & (nb >= 8u ? ~0ull : ((1ull << (nb * 8u)) - 1ull));
}

__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}

// `&[u8]` equality / ordering of Rust (lexicographic on unsigned bytes, then length)
__device__ __forceinline__ bool bytes_eq(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb) {
  if (la != lb) return false;
  for (int64_t k = 0; k < la; k += 8) {
    const uint32_t nb = (uint32_t)((la - k) < 8 ? (la - k) : 8);
    if (ld_upto8(a + k, nb) != ld_upto8(b + k, nb)) return false;
  }
  return true;
}
__device__ __forceinline__ bool bytes_lt(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb) {
  const int64_t n = la < lb ? la : lb;
  for (int64_t k = 0; k < n; k += 8) {
    const uint32_t nb = (uint32_t)((n - k) < 8 ? (n - k) : 8);
    const uint64_t x = ld_upto8(a + k, nb), y = ld_upto8(b + k, nb);
    if (x != y) return bswap64(x) < bswap64(y);  // first differing byte decides
  }
  return la < lb;
}

struct BytesItem { const uint8_t *p; int64_t len; };
__device__ __forceinline__ BytesItem bytes_item(const BytesOperand &s, int64_t i) {
  const int64_t b = ld_offset(s.offs, s.ob, i), e = ld_offset(s.offs, s.ob, i + 1);
  return BytesItem{s.data + b, e - b};
}

// ---- views (arrow-data/src/byte_view.rs): x = length, y = prefix / inline[0..4), z, w = inline[4..12) or (buffer index, offset)
__device__ __forceinline__ BytesItem view_item(const ViewOperand &s, const uint4 &v, const uint4 *slot) {
  if (v.x <= 12u) return BytesItem{reinterpret_cast<const uint8_t *>(slot) + 4, (int64_t)v.x};
  return BytesItem{s.buffers[v.z] + v.w, (int64_t)v.x};
}
__device__ __forceinline__ bool view_bits_eq(const uint4 &a, const uint4 &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
// GenericByteViewArray::inline_key_fast (byte_view_array.rs:872-874): (raw.swap_bytes() << 32) | len, compared as u128
__device__ __forceinline__ bool inline_key_lt(const uint4 &a, const uint4 &b) {
  // the key's 128 bits, most significant first: inline bytes 0..11 in order (big endian), then the length
  const uint32_t ak[4] = {__byte_perm(a.y, 0, 0x0123), __byte_perm(a.z, 0, 0x0123), __byte_perm(a.w, 0, 0x0123), a.x};
  const uint32_t bk[4] = {__byte_perm(b.y, 0, 0x0123), __byte_perm(b.z, 0, 0x0123), __byte_perm(b.w, 0, 0x0123), b.x};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (ak[k] != bk[k]) return ak[k] < bk[k];
  return false;
}
__device__ __forceinline__ bool view_is_eq(const ViewOperand &L, const uint4 &l, const uint4 *lslot, const ViewOperand &R, const uint4 &r,
                                           const uint4 *rslot) {  // cmp.rs:810-862
  if (L.n_buffers == 0 && R.n_buffers == 0) return view_bits_eq(l, r);
  if (view_bits_eq(l, r) && l.x <= 12u) return true;
  if (l.x != r.x) return false;
  if (l.x == 0u) return true;
  if (l.y != r.y) return false;
  if (l.x <= 12u) return false;
  const BytesItem a = view_item(L, l, lslot), b = view_item(R, r, rslot);
  return bytes_eq(a.p, a.len, b.p, b.len);
}
__device__ __forceinline__ bool view_is_lt(const ViewOperand &L, const uint4 &l, const uint4 *lslot, const ViewOperand &R, const uint4 &r,
                                           const uint4 *rslot) {  // cmp.rs:864-893
  if (L.n_buffers == 0 && R.n_buffers == 0) return inline_key_lt(l, r);
  if (l.x <= 12u && r.x <= 12u) return inline_key_lt(l, r);
  if (l.y != r.y) return __byte_perm(l.y, 0, 0x0123) < __byte_perm(r.y, 0, 0x0123);
  const BytesItem a = view_item(L, l, lslot), b = view_item(R, r, rslot);
  return bytes_lt(a.p, a.len, b.p, b.len);
}

// lane 0 of a 32-row group turns the ballot into the value / validity words (the word-level part of compare_op)
__device__ __forceinline__ void finish_group(const RowCmpCommon &p, int64_t row0, uint32_t v, int lane, unsigned &valid_cnt) {
  if (lane != 0) return;
  const int64_t left = p.n - row0;
  const uint32_t m = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
  if (p.neg) v = ~v;
  v &= m;
  uint32_t l = p.a_null_scalar ? 0u : m, r = p.b_null_scalar ? 0u : m;
  if (p.av) l &= ld_bits32(p.av, p.aoff + row0, p.aoff + p.n);
  if (p.bv) r &= ld_bits32(p.bv, p.boff + row0, p.boff + p.n);
  if (p.fold == SFOLD_DISTINCT) v = (l ^ r) | (l & r & v);                  // cmp.rs:331
  else if (p.fold == SFOLD_NOT_DISTINCT) v = (~(l | r) & m) | (l & r & v);  // cmp.rs:341
  p.out_bits[row0 >> 5] = v;
  if (p.out_valid) {
    p.out_valid[row0 >> 5] = l & r;
    valid_cnt += __popc(l & r);
  }
}

constexpr int ROWS_PER_LANE = 4;

__global__ void __launch_bounds__(256) k_cmp_bytes(const RowCmpCommon p, const BytesOperand A, const BytesOperand B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (p.n + 31) >> 5;
  unsigned valid_cnt = 0;
  BytesItem sa{nullptr, 0}, sb{nullptr, 0};
  if (p.a_scalar) sa = bytes_item(A, 0);
  if (p.b_scalar) sb = bytes_item(B, 0);
  for (int64_t g0 = warp * ROWS_PER_LANE; g0 < groups; g0 += nwarps * ROWS_PER_LANE) {
    BytesItem ia[ROWS_PER_LANE], ib[ROWS_PER_LANE];
#pragma unroll
    for (int k = 0; k < ROWS_PER_LANE; ++k) {  // the offset loads of 4 rows in flight
      const int64_t i = (g0 + k) * 32 + lane;
      const bool live = i < p.n;
      ia[k] = p.a_scalar ? sa : (live ? bytes_item(A, i) : BytesItem{nullptr, 0});
      ib[k] = p.b_scalar ? sb : (live ? bytes_item(B, i) : BytesItem{nullptr, 0});
    }
#pragma unroll
    for (int k = 0; k < ROWS_PER_LANE; ++k) {
      const int64_t row0 = (g0 + k) * 32;
      if (row0 >= p.n) break;  // warp-uniform
      const bool live = row0 + lane < p.n;
      bool r = false;
      if (live) r = p.lt ? bytes_lt(ia[k].p, ia[k].len, ib[k].p, ib[k].len) : bytes_eq(ia[k].p, ia[k].len, ib[k].p, ib[k].len);
      finish_group(p, row0, __ballot_sync(ACU_FULL_MASK, r), lane, valid_cnt);
    }
  }
  if (p.out_valid && lane == 0 && valid_cnt) atomicAdd(p.res + RES_COUNT, (unsigned long long)valid_cnt);
}

__global__ void __launch_bounds__(256) k_cmp_views(const RowCmpCommon p, const ViewOperand A, const ViewOperand B) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (p.n + 31) >> 5;
  unsigned valid_cnt = 0;
  uint4 sa = make_uint4(0, 0, 0, 0), sb = make_uint4(0, 0, 0, 0);
  if (p.a_scalar) sa = __ldg(A.views);
  if (p.b_scalar) sb = __ldg(B.views);
  for (int64_t g0 = warp * ROWS_PER_LANE; g0 < groups; g0 += nwarps * ROWS_PER_LANE) {
    uint4 va[ROWS_PER_LANE], vb[ROWS_PER_LANE];
#pragma unroll
    for (int k = 0; k < ROWS_PER_LANE; ++k) {
      const int64_t i = (g0 + k) * 32 + lane;
      const bool live = i < p.n;
      va[k] = p.a_scalar ? sa : (live ? ld_stream16(A.views + i) : make_uint4(0, 0, 0, 0));
      vb[k] = p.b_scalar ? sb : (live ? ld_stream16(B.views + i) : make_uint4(0, 0, 0, 0));
    }
#pragma unroll
    for (int k = 0; k < ROWS_PER_LANE; ++k) {
      const int64_t row0 = (g0 + k) * 32;
      if (row0 >= p.n) break;
      const int64_t i = row0 + lane;
      bool r = false;
      if (i < p.n) {
        const uint4 *as = A.views + (p.a_scalar ? 0 : i), *bs = B.views + (p.b_scalar ? 0 : i);
        r = p.lt ? view_is_lt(A, va[k], as, B, vb[k], bs) : view_is_eq(A, va[k], as, B, vb[k], bs);
      }
      finish_group(p, row0, __ballot_sync(ACU_FULL_MASK, r), lane, valid_cnt);
    }
  }
  if (p.out_valid && lane == 0 && valid_cnt) atomicAdd(p.res + RES_COUNT, (unsigned long long)valid_cnt);
}

// eq_inline_scalar (cmp.rs:405-435): (view as u64 & significant) == needle at every slot; 8 rows per lane in flight
__global__ void __launch_bounds__(256) k_view_eq_inline(const uint4 *__restrict__ views, int64_t n, uint64_t significant, uint64_t needle, int neg,
                                                        uint32_t *__restrict__ out_bits) {
  constexpr int U = 8;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (n + 31) >> 5;
  for (int64_t g0 = warp * U; g0 < groups; g0 += nwarps * U) {
    uint64_t lo[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t i = (g0 + k) * 32 + lane;
      lo[k] = i < n ? ld_stream8(views + i) : ~needle;  // the low 64 bits of the view: length + prefix
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t row0 = (g0 + k) * 32;
      if (row0 >= n) break;
      uint32_t v = __ballot_sync(ACU_FULL_MASK, (lo[k] & significant) == needle);
      if (lane == 0) {
        const int64_t left = n - row0;
        const uint32_t m = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
        if (neg) v = ~v;
        out_bits[row0 >> 5] = v & m;
      }
    }
  }
}

// ---- host side: compare_op (cmp.rs:220-382), the part that does not depend on the operand kind -----------------------
struct CmpPlan {
  RowCmpCommon p;
  bool swap;
  bool done;  // the result was produced without a comparison kernel (all-null)
};

acu_status new_null_bool(acu_ctx *ctx, int64_t len, acu_array_out *out) {  // BooleanArray::new_null(len)
  ACU_CUDA(ctx, cudaMemsetAsync(out->values, 0, acu_bitmap_bytes(len), ctx->stream));
  ACU_CUDA(ctx, cudaMemsetAsync(out->validity, 0, acu_bitmap_bytes(len), ctx->stream));
  ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->has_validity = 1;
  out->null_count = len;
  return ACU_OK;
}

acu_status cmp_prepare(acu_ctx *ctx, acu_cmp_op op, const acu_array *l, const acu_array *r, acu_array_out *out, CmpPlan *plan) {
  plan->done = false;
  const bool ls = l->is_scalar != 0, rs = r->is_scalar != 0;
  if (l->len != r->len && !ls && !rs)  // cmp.rs:228-232
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Cannot compare arrays of different lengths, got %lld vs %lld",
                    (long long)l->len, (long long)r->len);
  const int64_t len = ls ? r->len : l->len;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  if (len == 0) { plan->done = true; return ACU_OK; }
  acu_status st;
  const int64_t lnc = acu_resolve_null_count(ctx, l, &st);
  ACU_TRY(st);
  const int64_t rnc = acu_resolve_null_count(ctx, r, &st);
  ACU_TRY(st);
  const bool ln = lnc > 0, rn = rnc > 0;
  const bool fold = op == ACU_DISTINCT || op == ACU_NOT_DISTINCT;
  const bool l_null_scalar = ls && ln, r_null_scalar = rs && rn;
  if (!fold && (l_null_scalar || r_null_scalar) && !(ls && rs)) {  // cmp.rs:353, :364
    plan->done = true;
    return new_null_bool(ctx, len, out);
  }
  RowCmpCommon &p = plan->p;
  p = RowCmpCommon{};
  p.n = len;
  p.res = ctx->d_res;
  p.out_bits = static_cast<uint32_t *>(out->values);
  plan->swap = (op == ACU_GT || op == ACU_LT_EQ);  // cmp.rs:481-488
  const bool xs = plan->swap ? rs : ls, ys = plan->swap ? ls : rs;
  const bool xn = plan->swap ? rn : ln, yn = plan->swap ? ln : rn;
  const acu_array *x = plan->swap ? r : l, *y = plan->swap ? l : r;
  p.a_scalar = xs && !(xs && ys);
  p.b_scalar = ys && !(xs && ys);
  p.neg = (op == ACU_NEQ || op == ACU_DISTINCT || op == ACU_LT_EQ || op == ACU_GT_EQ);
  p.fold = op == ACU_DISTINCT ? SFOLD_DISTINCT : op == ACU_NOT_DISTINCT ? SFOLD_NOT_DISTINCT : SFOLD_NONE;
  p.lt = !(op == ACU_EQ || op == ACU_NEQ || fold);
  if (xn) { if (p.a_scalar) p.a_null_scalar = 1; else { p.av = x->validity; p.aoff = x->validity_offset; } }
  if (yn) { if (p.b_scalar) p.b_null_scalar = 1; else { p.bv = y->validity; p.boff = y->validity_offset; } }
  if (!fold && (xn || yn)) p.out_valid = reinterpret_cast<uint32_t *>(out->validity);
  return ACU_OK;
}

acu_status cmp_finish(acu_ctx *ctx, const CmpPlan &plan, acu_array_out *out) {
  ACU_TRY(acu_res_fetch(ctx));
  if (plan.p.out_valid) {
    out->has_validity = 1;
    out->null_count = plan.p.n - (int64_t)ctx->h_res[RES_COUNT];
  }
  return ACU_OK;
}

int cmp_grid(acu_ctx *ctx, int64_t n, int rows_per_lane) {
  const int64_t groups = (n + 31) / 32;
  return acu_grid(ctx, ((groups + rows_per_lane - 1) / rows_per_lane + 7) / 8, 16);
}

}  // namespace

extern "C" acu_status acu_cmp_bytes(acu_ctx *ctx, int32_t offset_bytes, acu_cmp_op op, const acu_bytes_array *l, const acu_bytes_array *r,
                                    acu_array_out *out) {
  ACU_ENTER(ctx);
  if (offset_bytes != 4 && offset_bytes != 8) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "offset width must be 4 or 8");
  CmpPlan plan;
  ACU_TRY(cmp_prepare(ctx, op, &l->nulls, &r->nulls, out, &plan));
  if (plan.done) return ACU_OK;
  const acu_bytes_array *x = plan.swap ? r : l, *y = plan.swap ? l : r;
  const BytesOperand A{x->offsets, x->data, offset_bytes}, B{y->offsets, y->data, offset_bytes};
  ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, k_cmp_bytes, cmp_grid(ctx, plan.p.n, ROWS_PER_LANE), 256, 0, plan.p, A, B);
  return cmp_finish(ctx, plan, out);
}

extern "C" acu_status acu_cmp_byte_view(acu_ctx *ctx, acu_cmp_op op, const acu_view_array *l, const acu_view_array *r, acu_array_out *out) {
  ACU_ENTER(ctx);
  const bool ls = l->nulls.is_scalar != 0, rs = r->nulls.is_scalar != 0;
  // eq_inline_scalar (cmp.rs:282-300): == / != of an array against a non-null constant of <= 4 bytes
  if ((op == ACU_EQ || op == ACU_NEQ) && ls != rs) {
    const acu_view_array *arr = ls ? r : l, *sc = ls ? l : r;
    acu_status st;
    const int64_t snc = sc->nulls.len >= 1 ? acu_resolve_null_count(ctx, &sc->nulls, &st) : 1;
    if (sc->nulls.len >= 1) ACU_TRY(st);
    if (sc->nulls.len >= 1 && snc == 0) {
      uint64_t low = 0;
      ACU_CUDA(ctx, cudaMemcpyAsync(&low, sc->views, 8, cudaMemcpyDeviceToHost, ctx->stream));
      ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      const uint32_t needle_len = (uint32_t)low;
      if (needle_len <= 4) {
        const uint64_t significant = ~0ull >> (32 - needle_len * 8);
        const int64_t n = arr->nulls.len;
        out->len = n;
        out->has_validity = 0;
        out->null_count = 0;
        if (n == 0) return ACU_OK;
        const int64_t anc = acu_resolve_null_count(ctx, &arr->nulls, &st);
        ACU_TRY(st);
        ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, k_view_eq_inline, cmp_grid(ctx, n, 8), 256, 0, static_cast<const uint4 *>(arr->views), n, significant,
                         low & significant, op == ACU_NEQ ? 1 : 0, static_cast<uint32_t *>(out->values));
        if (arr->nulls.validity && anc > 0) {  // nulls = values_nulls.filter(null_count > 0)
          ACU_TRY(acu_bitmap_and_launch(ctx, arr->nulls.validity, arr->nulls.validity_offset, nullptr, 0, n,
                                        reinterpret_cast<uint64_t *>(out->validity), false));
          out->has_validity = 1;
          out->null_count = anc;
        }
        ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        acu_kstats_drain(ctx);
        return ACU_OK;
      }
    }
  }
  CmpPlan plan;
  ACU_TRY(cmp_prepare(ctx, op, &l->nulls, &r->nulls, out, &plan));
  if (plan.done) return ACU_OK;
  const acu_view_array *x = plan.swap ? r : l, *y = plan.swap ? l : r;
  // the data-buffer pointer tables go to the device (scratch): [x buffers][y buffers]
  const int nx = x->n_buffers, ny = y->n_buffers;
  const uint8_t **table = nullptr;
  if (nx + ny > 0) {
    void *scratch;
    ACU_TRY(acu_scratch(ctx, (size_t)(nx + ny) * sizeof(void *), &scratch));
    table = static_cast<const uint8_t **>(scratch);
    if (nx) ACU_CUDA(ctx, cudaMemcpyAsync(table, x->buffers, (size_t)nx * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
    if (ny) ACU_CUDA(ctx, cudaMemcpyAsync(table + nx, y->buffers, (size_t)ny * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
  }
  const ViewOperand A{static_cast<const uint4 *>(x->views), table, nx}, B{static_cast<const uint4 *>(y->views), table ? table + nx : nullptr, ny};
  ACU_TRY(acu_res_reset(ctx));
  ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, k_cmp_views, cmp_grid(ctx, plan.p.n, ROWS_PER_LANE), 256, 0, plan.p, A, B);
  return cmp_finish(ctx, plan, out);
}
