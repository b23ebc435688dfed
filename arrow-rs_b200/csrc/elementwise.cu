// elementwise.cu — numeric::{add,sub,mul,div,rem,neg}(_wrapping), cmp::*, cast (numeric).
//
// Reference: arrow-arith/src/numeric.rs:36-374, arrow-arith/src/arity.rs:104-135,254-299,
// arrow-array/src/array/primitive_array.rs:916-1103, arrow-array/src/arithmetic.rs:148-437,
// arrow-ord/src/cmp.rs:220-648, arrow-cast/src/cast/mod.rs:2550-2614.
//
// All kernels are single-pass and HBM-bound: each input byte is read once, each output
// byte written once, and the validity AND + popcount ride along in the same pass.
// Work unit = a "strip" of R rows handled by one warp: R = max(64, 32*EPL) where EPL =
// elements per 128-bit lane access. Lane l owns elements [l*EPL, (l+1)*EPL) of each
// 32*EPL-row load, so every warp access is a fully coalesced 512-B (EPL>1) request, and a
// strip always spans whole 64-bit validity words (lanes 0..R/64-1 own one word each).
#include <stdio.h>

#include <limits>
#include <type_traits>

#include "bitmap.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// 128-bit packs
// ---------------------------------------------------------------------------------------
template <class T, int EPL> struct alignas(EPL * sizeof(T) == 16 ? 16 : sizeof(T)) Pack { T v[EPL]; };

template <class T, int EPL>
__device__ __forceinline__ Pack<T, EPL> pack_load(const T *p) {
  Pack<T, EPL> r;
  if constexpr (EPL * sizeof(T) == 16) {
    uint4 x = ld_stream16(p);
    r = *reinterpret_cast<Pack<T, EPL> *>(&x);
  } else {
    static_assert(EPL == 1, "scalar path");
    r.v[0] = __ldg(p);
  }
  return r;
}
template <class T, int EPL>
__device__ __forceinline__ void pack_store(T *p, Pack<T, EPL> r) {
  if constexpr (EPL * sizeof(T) == 16) {
    st_stream16(p, *reinterpret_cast<uint4 *>(&r));
  } else {
    *p = r.v[0];
  }
}

__device__ __forceinline__ uint64_t ones_to(int64_t row, int64_t n) {  // bits [row,row+64) ∩ [0,n)
  int64_t k = n - row;
  return k >= 64 ? ~0ull : (k <= 0 ? 0ull : ((~0ull) >> (64 - k)));
}

// ---------------------------------------------------------------------------------------
// Per-element arithmetic (ArrowNativeTypeOp, arithmetic.rs:148-437)
// ---------------------------------------------------------------------------------------
enum { CLS_WRAP = 0, CLS_CHECKED = 1, CLS_DIVREM = 2 };
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_REM = 4, OP_NEG = 5 };

template <class T> struct is_fp { static constexpr bool value = std::is_floating_point<T>::value; };

__device__ __forceinline__ double fp_add(double a, double b) { return __dadd_rn(a, b); }  // never contracted
__device__ __forceinline__ double fp_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fp_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fp_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double fp_rem(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float fp_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fp_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fp_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fp_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fp_rem(float a, float b) { return fmodf(a, b); }

// returns true when the reference would return Err at this element
template <class T, int CLS>
__device__ __forceinline__ bool apply_op(int op, T l, T r, T &o) {
  if constexpr (is_fp<T>::value) {
    if constexpr (CLS == CLS_DIVREM) {
      o = (op == OP_DIV) ? fp_div(l, r) : fp_rem(l, r);
    } else {
      o = (op == OP_ADD) ? fp_add(l, r) : (op == OP_SUB) ? fp_sub(l, r) : (op == OP_MUL) ? fp_mul(l, r) : -r;
    }
    return false;
  } else {
    using U = typename std::make_unsigned<T>::type;
    constexpr bool SIGNED = std::is_signed<T>::value;
    if constexpr (CLS == CLS_WRAP) {
      U a = (U)l, b = (U)r;
      o = (T)((op == OP_ADD) ? (U)(a + b) : (op == OP_SUB) ? (U)(a - b) : (op == OP_MUL) ? (U)(a * b) : (U)((U)0 - b));
      return false;
    } else if constexpr (CLS == CLS_CHECKED) {
      if constexpr (sizeof(T) < 8) {
        using W = typename std::conditional<SIGNED, int64_t, uint64_t>::type;
        W a = (W)l, b = (W)r;
        W w = (op == OP_ADD) ? a + b : (op == OP_SUB) ? a - b : (op == OP_MUL) ? a * b : (W)0 - b;
        o = (T)w;
        if constexpr (SIGNED) return w < (W)std::numeric_limits<T>::min() || w > (W)std::numeric_limits<T>::max();
        else return (op == OP_SUB) ? (a < b) : (op == OP_NEG ? false : w > (W)std::numeric_limits<T>::max());
      } else if constexpr (SIGNED) {
        int64_t a = l, b = r;
        if (op == OP_NEG) { a = 0; }
        if (op == OP_ADD) {
          int64_t s = (int64_t)((uint64_t)a + (uint64_t)b);
          o = s;
          return ((a ^ s) & (b ^ s)) < 0;
        } else if (op == OP_MUL) {
          int64_t lo = (int64_t)((uint64_t)a * (uint64_t)b);
          int64_t hi = __mul64hi(a, b);
          o = lo;
          return hi != (lo >> 63);
        } else {  // SUB / NEG
          int64_t s = (int64_t)((uint64_t)a - (uint64_t)b);
          o = s;
          return ((a ^ b) & (a ^ s)) < 0;
        }
      } else {
        uint64_t a = l, b = r;
        if (op == OP_ADD) { o = a + b; return o < a; }
        if (op == OP_MUL) { o = a * b; return __umul64hi(a, b) != 0; }
        if (op == OP_NEG) { o = (uint64_t)0 - b; return false; }
        o = a - b;
        return a < b;
      }
    } else {  // CLS_DIVREM: div_checked (arithmetic.rs:204-215) / rem (numeric.rs:345-351)
      o = 0;
      if (r == 0) return true;
      if constexpr (SIGNED) {
        if (r == (T)-1) {
          if (op == OP_DIV) {
            if (l == std::numeric_limits<T>::min()) return true;
            o = (T)(-l);
          }
          return false;  // rem: wrapping_rem(x, -1) == 0
        }
      }
      o = (op == OP_DIV) ? (T)(l / r) : (T)(l % r);
      return false;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Binary / unary arithmetic kernel
// ---------------------------------------------------------------------------------------
template <class T>
struct ArithParams {
  const T *a, *b;
  T *out;
  int64_t n;
  const uint8_t *av, *bv;  // input validity bitmaps taking part in the union (or NULL)
  int64_t aoff, boff;
  uint64_t *out_valid;     // NULL => no NullBuffer in the result
  unsigned long long *res;
  int op;
  int a_scalar, b_scalar;
  int zero_nulls;          // try_binary / try_unary: zero under nulls, op only at valid slots
};

// Steady state: a warp owns "super-groups" of 2048 rows = 32 validity words (lane l <-> word
// l: ONE coalesced 256-B bitmap access per operand per super-group), processed as groups of
// U strips whose loads are all issued before any use. The ragged remainder (< 2048 rows) is
// finished element-wise by warp 0 so that bounds-checked code stays out of the streaming loop.
template <class T, int CLS, int EPL>
__global__ void __launch_bounds__(256, (CLS == CLS_WRAP ? 4 : 3)) k_arith(const ArithParams<T> p) {
  constexpr int R = (32 * EPL > 64) ? 32 * EPL : 64;  // rows per strip
  constexpr int LPS = R / (32 * EPL);                 // loads per lane per strip
  constexpr int U = (LPS >= 2) ? 2 : 4;               // strips in flight per warp
  constexpr int GROUP = U * R;                        // rows per group
  constexpr int SG = 2048;                            // rows per super-group
  constexpr int GPS = SG / GROUP;                     // groups per super-group
  constexpr bool fallible = (CLS != CLS_WRAP) && !is_fp<T>::value;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = p.n;
  const int64_t sgroups = n / SG;
  const bool has_valid = p.out_valid != nullptr;
  T sa = T(), sb = T();
  if (p.a_scalar) sa = __ldg(p.a);
  if (p.b_scalar) sb = __ldg(p.b);
  unsigned valid_cnt = 0;
  unsigned long long err = ~0ull;

  for (int64_t sg = warp; sg < sgroups; sg += nwarps) {
    const int64_t sbase = sg * SG;
    uint64_t vw = ~0ull;  // lane l owns validity word l of the super-group
    if (has_valid) {
      const int64_t row = sbase + lane * 64;
      if (p.av) vw &= ld_bits64(p.av, p.aoff + row, p.aoff + n);
      if (p.bv) vw &= ld_bits64(p.bv, p.boff + row, p.boff + n);
      p.out_valid[row >> 6] = vw;
      valid_cnt += __popcll(vw);
    }
#pragma unroll 1
    for (int gi = 0; gi < GPS; ++gi) {
      const int64_t base = sbase + gi * GROUP;
      const T *__restrict__ pa = p.a + base + lane * EPL;
      const T *__restrict__ pb = p.b + base + lane * EPL;
      Pack<T, EPL> va[U * LPS], vb[U * LPS];
      // ---- every load of the group is issued before any use (memory-level parallelism) ----
#pragma unroll
      for (int k = 0; k < U * LPS; ++k) {
        if (!p.a_scalar) va[k] = pack_load<T, EPL>(pa + k * 32 * EPL);
        if (!p.b_scalar) vb[k] = pack_load<T, EPL>(pb + k * 32 * EPL);
      }
      // ---- compute + store ----
      T *__restrict__ po = p.out + base + lane * EPL;
#pragma unroll
      for (int k = 0; k < U * LPS; ++k) {
        uint32_t bits = ~0u;
        if (fallible && p.zero_nulls) {
          const int pos = gi * GROUP + k * 32 * EPL + lane * EPL;  // row inside the super-group
          const uint64_t w = __shfl_sync(ACU_FULL_MASK, vw, pos >> 6);
          bits = (uint32_t)(w >> (pos & 63));
        }
        Pack<T, EPL> o;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const T l = p.a_scalar ? sa : va[k].v[e];
          const T r = p.b_scalar ? sb : vb[k].v[e];
          T x;
          const bool bad = apply_op<T, CLS>(p.op, l, r, x);
          if (fallible) {
            if (!((bits >> e) & 1u)) x = T();
            else if (bad) {
              const unsigned long long i = (unsigned long long)(base + k * 32 * EPL + lane * EPL + e);
              err = i < err ? i : err;
            }
          }
          o.v[e] = x;
        }
        pack_store<T, EPL>(po + k * 32 * EPL, o);
      }
    }
  }

  // ---- ragged remainder: 64-row strips, lane owns rows l and l+32 ----
  if (warp == 0) {
    for (int64_t row = sgroups * SG; row < n; row += 64) {
      uint64_t vw = ones_to(row, n);
      if (has_valid) {
        if (p.av) vw &= ld_bits64(p.av, p.aoff + row, p.aoff + n);
        if (p.bv) vw &= ld_bits64(p.bv, p.boff + row, p.boff + n);
        if (lane == 0) { p.out_valid[row >> 6] = vw; valid_cnt += __popcll(vw); }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t i = row + h * 32 + lane;
        if (i >= n) continue;
        const T l = p.a_scalar ? sa : __ldg(p.a + i);
        const T r = p.b_scalar ? sb : __ldg(p.b + i);
        T x;
        const bool bad = apply_op<T, CLS>(p.op, l, r, x);
        if (fallible) {
          const bool live = !p.zero_nulls || ((vw >> (h * 32 + lane)) & 1ull);
          if (!live) x = T();
          else if (bad) err = (unsigned long long)i < err ? (unsigned long long)i : err;
        }
        p.out[i] = x;
      }
    }
  }
  if (has_valid) {
    valid_cnt = warp_sum(valid_cnt);
    if (lane == 0 && valid_cnt) atomicAdd(p.res + RES_COUNT, (unsigned long long)valid_cnt);
  }
  if (fallible && err != ~0ull) atomicMin(p.res + RES_ERR_INDEX, err);
}

template <class T, int CLS>
acu_status launch_arith(acu_ctx *ctx, const ArithParams<T> &p) {
  constexpr int EPLV = 16 / sizeof(T);
  bool aligned = ((uintptr_t)p.out % 16 == 0) && (p.a_scalar || (uintptr_t)p.a % 16 == 0) &&
                 (p.b_scalar || (uintptr_t)p.b % 16 == 0);
  if (aligned) {
    int64_t blocks = (p.n / 2048 + 1 + 7) / 8;  // 8 warps per CTA, one 2048-row super-group per warp step
    ACU_LAUNCH_TIMED(ctx, ACU_K_ARITH, (k_arith<T, CLS, EPLV>), acu_wave_grid(ctx, k_arith<T, CLS, EPLV>, 256, 0, blocks), 256, 0, p);
  } else {
    int64_t blocks = (p.n / 2048 + 1 + 7) / 8;
    ACU_LAUNCH_TIMED(ctx, ACU_K_ARITH, (k_arith<T, CLS, 1>), acu_wave_grid(ctx, k_arith<T, CLS, 1>, 256, 0, blocks), 256, 0, p);
  }
  return ACU_OK;
}

const char *op_symbol(acu_arith_op op) {  // numeric.rs:192-202
  switch (op) {
    case ACU_ADD_WRAPPING: case ACU_ADD: return "+";
    case ACU_SUB_WRAPPING: case ACU_SUB: return "-";
    case ACU_MUL_WRAPPING: case ACU_MUL: return "*";
    case ACU_DIV: return "/";
    default: return "%";
  }
}

template <class T> void fmt_native(char *buf, size_t n, T v) {  // Rust {:?} of integers
  if constexpr (std::is_floating_point<T>::value) snprintf(buf, n, "%.17g", (double)v);
  else if constexpr (std::is_signed<T>::value) snprintf(buf, n, "%lld", (long long)v);
  else snprintf(buf, n, "%llu", (unsigned long long)v);
}
template <class T> uint64_t bits_of(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }

// Fetch the operands at the lowest failing row and rebuild the reference's error.
template <class T>
acu_status arith_error(acu_ctx *ctx, acu_arith_op op, bool is_neg, const acu_array *a, const acu_array *b, int64_t idx) {
  T l = T(), r = T();
  if (a) ACU_CUDA(ctx, cudaMemcpyAsync(&l, static_cast<const T *>(a->values) + (a->is_scalar ? 0 : idx), sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  ACU_CUDA(ctx, cudaMemcpyAsync(&r, static_cast<const T *>(b->values) + (b->is_scalar ? 0 : idx), sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  char ls[40], rs[40];
  fmt_native(ls, sizeof ls, l);
  fmt_native(rs, sizeof rs, r);
  if (is_neg)
    return acu_fail(ctx, ACU_ERR_ARITHMETIC_OVERFLOW, idx, bits_of(r), 0, 0, "Overflow happened on: - %s", rs);
  if ((op == ACU_DIV || op == ACU_REM) && r == T())
    return acu_fail(ctx, ACU_ERR_DIVIDE_BY_ZERO, idx, bits_of(l), bits_of(r), 0, "Divide by zero error");
  return acu_fail(ctx, ACU_ERR_ARITHMETIC_OVERFLOW, idx, bits_of(l), bits_of(r), 0,
                  "Overflow happened on: %s %s %s", ls, op_symbol(op), rs);
}

acu_status set_new_null(acu_ctx *ctx, int64_t len, size_t value_bytes, acu_array_out *out) {
  // PrimitiveArray::new_null / BooleanArray::new_null: zeroed values, all-null bitmap
  if (value_bytes) ACU_CUDA(ctx, cudaMemsetAsync(out->values, 0, value_bytes, ctx->stream));
  if (len) ACU_CUDA(ctx, cudaMemsetAsync(out->validity, 0, acu_bitmap_bytes(len), ctx->stream));
  ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->len = len;
  out->has_validity = 1;
  out->null_count = len;
  return ACU_OK;
}

template <class T>
acu_status arith_typed(acu_ctx *ctx, acu_arith_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  const bool checked = !is_fp<T>::value && (op == ACU_ADD || op == ACU_SUB || op == ACU_MUL || op == ACU_DIV || op == ACU_REM);
  const bool a_s = a->is_scalar != 0, b_s = b->is_scalar != 0;
  acu_status st;
  ArithParams<T> p{};
  p.a = static_cast<const T *>(a->values);
  p.b = static_cast<const T *>(b->values);
  p.out = static_cast<T *>(out->values);
  p.res = ctx->d_res;
  p.a_scalar = a_s && !b_s;
  p.b_scalar = b_s && !a_s;
  switch (op) {
    case ACU_ADD_WRAPPING: case ACU_ADD: p.op = OP_ADD; break;
    case ACU_SUB_WRAPPING: case ACU_SUB: p.op = OP_SUB; break;
    case ACU_MUL_WRAPPING: case ACU_MUL: p.op = OP_MUL; break;
    case ACU_DIV: p.op = OP_DIV; break;
    default: p.op = OP_REM; break;
  }
  out->has_validity = 0;
  out->null_count = 0;
  int64_t len;
  if (a_s != b_s) {  // op!/try_op! scalar arms (numeric.rs:278-317)
    const acu_array *s = a_s ? a : b, *arr = a_s ? b : a;
    len = arr->len;
    int64_t snc = acu_resolve_null_count(ctx, s, &st);
    ACU_TRY(st);
    if (snc != 0) return set_new_null(ctx, len, (size_t)len * sizeof(T), out);
    out->len = len;
    if (len == 0) { out->has_validity = arr->validity != nullptr; return ACU_OK; }
    if (arr->validity) {  // nulls().cloned()
      (a_s ? p.bv : p.av) = arr->validity;
      (a_s ? p.boff : p.aoff) = arr->validity_offset;
      p.out_valid = reinterpret_cast<uint64_t *>(out->validity);
      p.zero_nulls = checked;
    }
  } else {  // binary / try_binary (arity.rs:104-135, :254-299)
    if (a->len != b->len)
      return acu_fail(ctx, ACU_ERR_COMPUTE, -1, 0, 0, 0,
                      checked ? "Cannot perform a binary operation on arrays of different length"
                              : "Cannot perform binary operation on arrays of different length");
    len = a->len;
    out->len = len;
    if (len == 0) return ACU_OK;
    int64_t an = acu_resolve_null_count(ctx, a, &st);
    ACU_TRY(st);
    int64_t bn = acu_resolve_null_count(ctx, b, &st);
    ACU_TRY(st);
    // NullBuffer::union (null.rs:79-87)
    if (a->validity && b->validity && (an > 0 || bn > 0)) {
      p.av = a->validity; p.aoff = a->validity_offset;
      p.bv = b->validity; p.boff = b->validity_offset;
    } else if (a->validity && !b->validity && an > 0) {
      p.av = a->validity; p.aoff = a->validity_offset;
    } else if (b->validity && !a->validity && bn > 0) {
      p.bv = b->validity; p.boff = b->validity_offset;
    }
    if (p.av || p.bv) {
      p.out_valid = reinterpret_cast<uint64_t *>(out->validity);
      p.zero_nulls = checked;
    }
  }
  p.n = len;
  const int blk = acu_call_begin(ctx, &st);
  ACU_TRY(st);
  p.res = acu_dres(ctx, blk);
  if (is_fp<T>::value) {
    if (op == ACU_DIV || op == ACU_REM) ACU_TRY((launch_arith<T, CLS_DIVREM>(ctx, p)));
    else ACU_TRY((launch_arith<T, CLS_WRAP>(ctx, p)));
  } else if (op == ACU_DIV || op == ACU_REM) {
    ACU_TRY((launch_arith<T, CLS_DIVREM>(ctx, p)));
  } else if (checked) {
    ACU_TRY((launch_arith<T, CLS_CHECKED>(ctx, p)));
  } else {
    ACU_TRY((launch_arith<T, CLS_WRAP>(ctx, p)));
  }
  const acu_array ca = *a, cb = *b;  // the finaliser may run later (acu_results_fetch)
  const bool has_valid = p.out_valid != nullptr;
  return acu_call_end(ctx, blk, [ctx, op, checked, ca, cb, has_valid, len, out](const unsigned long long *h) -> acu_status {
    if (checked && h[RES_ERR_INDEX] != ~0ull) return arith_error<T>(ctx, op, false, &ca, &cb, (int64_t)h[RES_ERR_INDEX]);
    if (has_valid) {
      out->has_validity = 1;
      out->null_count = len - (int64_t)h[RES_COUNT];
    }
    return ACU_OK;
  });
}

template <class T>
acu_status neg_typed(acu_ctx *ctx, int32_t checked_in, const acu_array *a, acu_array_out *out) {
  const bool checked = checked_in && !is_fp<T>::value;
  int64_t len = a->len;
  out->len = len;
  out->has_validity = a->validity != nullptr;
  out->null_count = 0;
  if (len == 0) return ACU_OK;
  ArithParams<T> p{};
  p.a = static_cast<const T *>(a->values);  // ignored by OP_NEG (a_scalar => one broadcast load)
  p.b = static_cast<const T *>(a->values);
  p.out = static_cast<T *>(out->values);
  p.res = ctx->d_res;
  p.a_scalar = 1;
  p.op = OP_NEG;
  p.n = len;
  if (a->validity) {  // unary / try_unary: nulls().cloned()
    p.bv = a->validity;
    p.boff = a->validity_offset;
    p.out_valid = reinterpret_cast<uint64_t *>(out->validity);
    p.zero_nulls = checked;
  }
  ACU_TRY(acu_res_reset(ctx));
  if (checked) ACU_TRY((launch_arith<T, CLS_CHECKED>(ctx, p)));
  else ACU_TRY((launch_arith<T, CLS_WRAP>(ctx, p)));
  ACU_TRY(acu_res_fetch(ctx));
  if (checked && ctx->h_res[RES_ERR_INDEX] != ~0ull)
    return arith_error<T>(ctx, ACU_SUB, true, nullptr, a, (int64_t)ctx->h_res[RES_ERR_INDEX]);
  if (p.out_valid) out->null_count = len - (int64_t)ctx->h_res[RES_COUNT];
  return ACU_OK;
}

// ---------------------------------------------------------------------------------------
// cmp — one result bit per row (collect_bool, cmp.rs:580-611)
// ---------------------------------------------------------------------------------------
enum { FOLD_NONE = 0, FOLD_DISTINCT = 1, FOLD_NOT_DISTINCT = 2 };

template <class T>
struct CmpParams {
  const T *a, *b;
  int64_t n;
  const uint8_t *av, *bv;
  int64_t aoff, boff;
  int a_scalar, b_scalar;
  int a_null_scalar, b_null_scalar;  // scalar side whose single slot is null
  int neg, fold;
  uint64_t *out_bits, *out_valid;
  unsigned long long *res;
  // fused compare -> filter plan (acu_filter_plan_create_cmp): out_bits is the plan's mask and receives result & validity;
  // tile_count[t] = selected rows of 1024-row tile t (written by the streaming kernel; the tail's tiles are counted separately)
  int fuse;
  uint32_t *tile_count;
};

template <class T> __device__ __forceinline__ bool pred_eq(T l, T r) {
  if constexpr (sizeof(T) == 8 && is_fp<T>::value) return __double_as_longlong(l) == __double_as_longlong(r);
  else if constexpr (is_fp<T>::value) return __float_as_int(l) == __float_as_int(r);
  else return l == r;
}
template <class T> __device__ __forceinline__ bool pred_lt(T l, T r) {
  if constexpr (is_fp<T>::value) return total_key(l) < total_key(r);
  else return l < r;
}

// Lane l of a warp owns rows l and l+32 of each 64-row strip: two ballots give the low and
// high halves of the packed u64, and the loads are coalesced 32*sizeof(T)-byte requests.
template <class T, bool LT>
__global__ void __launch_bounds__(256) k_cmp(const CmpParams<T> p) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = p.n;
  const int64_t strips = (n + 63) >> 6;
  T sa = T(), sb = T();
  if (p.a_scalar) sa = __ldg(p.a);
  if (p.b_scalar) sb = __ldg(p.b);
  unsigned valid_cnt = 0;
  for (int64_t s0 = warp * U; s0 < strips; s0 += nwarps * U) {
    T va[U][2], vb[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t i = (s0 + u) * 64 + h * 32 + lane;
        va[u][h] = (!p.a_scalar && i < n) ? __ldg(p.a + i) : sa;
        vb[u][h] = (!p.b_scalar && i < n) ? __ldg(p.b + i) : sb;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = (s0 + u) * 64;
      if (row >= n) break;
      bool r0 = LT ? pred_lt(va[u][0], vb[u][0]) : pred_eq(va[u][0], vb[u][0]);
      bool r1 = LT ? pred_lt(va[u][1], vb[u][1]) : pred_eq(va[u][1], vb[u][1]);
      uint64_t v = (uint64_t)__ballot_sync(ACU_FULL_MASK, r0) | ((uint64_t)__ballot_sync(ACU_FULL_MASK, r1) << 32);
      if (p.neg) v = ~v;
      if (lane == 0) {
        const uint64_t m = ones_to(row, n);
        v &= m;
        uint64_t l = p.a_null_scalar ? 0ull : m, r = p.b_null_scalar ? 0ull : m;
        if (p.av) l &= ld_bits64(p.av, p.aoff + row, p.aoff + n);
        if (p.bv) r &= ld_bits64(p.bv, p.boff + row, p.boff + n);
        if (p.fold == FOLD_DISTINCT) v = (l ^ r) | (l & r & v);              // cmp.rs:331
        else if (p.fold == FOLD_NOT_DISTINCT) v = (~(l | r) & m) | (l & r & v);  // cmp.rs:341
        else if (p.fuse) v &= l & r;  // a null result selects nothing (filter.rs:167-171)
        p.out_bits[row >> 6] = v;
        if (p.out_valid) {
          p.out_valid[row >> 6] = l & r;
          valid_cnt += __popcll(l & r);
        }
      }
    }
  }
  if (p.out_valid && lane == 0 && valid_cnt) atomicAdd(p.res + RES_COUNT, (unsigned long long)valid_cnt);
}

// Streaming variant for 16-B aligned operands: a warp owns 2048-row super-groups = 32 result
// words. Every lane reads 128-bit vectors (EPL elements), 4 vector pairs in flight; a lane's
// EPL result bits sit at bit position lane*EPL of the load's bit string, so each 32-bit half
// of a result word is ONE warp OR-reduction (redux.sync) of the shifted groups. Lane l ends up
// owning result word l and validity word l: one coalesced 256-B store each per super-group.
// The ragged tail (< 2048 rows) is finished by k_cmp on offset pointers.
template <class T, bool LT>
__global__ void __launch_bounds__(256, 4) k_cmp_v2(const CmpParams<T> p, const int64_t sgroups) {
  constexpr int EPL = 16 / sizeof(T);
  constexpr int RPL = 32 * EPL;          // rows per warp-wide vector load
  constexpr int HPL = RPL / 32;          // 32-bit result halves per load (= EPL)
  constexpr int LOADS = 2048 / RPL;      // vector loads per super-group (32 / 16 / 8 / 4)
  constexpr int U = LOADS < 4 ? LOADS : 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n = p.n;
  T sa = T(), sb = T();
  if (p.a_scalar) sa = __ldg(p.a);
  if (p.b_scalar) sb = __ldg(p.b);
  unsigned valid_cnt = 0;
  const int grp_shift = (lane * EPL) & 31;   // where this lane's EPL bits land inside their 32-bit half
  const int grp_half = (lane * EPL) >> 5;    // which half of the load's bit string they belong to
  for (int64_t sg = warp; sg < sgroups; sg += nwarps) {
    const int64_t sbase = sg << 11;
    uint64_t lw = ~0ull, rw = ~0ull;  // lane-owned validity words
    if (p.a_null_scalar) lw = 0;
    if (p.b_null_scalar) rw = 0;
    const int64_t wrow = sbase + lane * 64;
    if (p.av) lw &= ld_bits64(p.av, p.aoff + wrow, p.aoff + n);
    if (p.bv) rw &= ld_bits64(p.bv, p.boff + wrow, p.boff + n);
    uint32_t my_lo = 0, my_hi = 0;    // lane-owned result word
#pragma unroll 1
    for (int l0 = 0; l0 < LOADS; l0 += U) {
      Pack<T, EPL> va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i0 = sbase + (int64_t)(l0 + u) * RPL + lane * EPL;
        if (!p.a_scalar) va[u] = pack_load<T, EPL>(p.a + i0);
        if (!p.b_scalar) vb[u] = pack_load<T, EPL>(p.b + i0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t x = 0;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const T l = p.a_scalar ? sa : va[u].v[e];
          const T r = p.b_scalar ? sb : vb[u].v[e];
          x |= (uint32_t)(LT ? pred_lt(l, r) : pred_eq(l, r)) << e;
        }
        x <<= grp_shift;
#pragma unroll
        for (int hh = 0; hh < HPL; ++hh) {
          const uint32_t half = __reduce_or_sync(ACU_FULL_MASK, grp_half == hh ? x : 0u);
          const int widx = ((l0 + u) * HPL + hh) >> 1;  // result word inside the super-group
          if (lane == widx) { if (hh & 1) my_hi = half; else my_lo = half; }
        }
      }
    }
    uint64_t v = (uint64_t)my_lo | ((uint64_t)my_hi << 32);
    if (p.neg) v = ~v;
    if (p.fold == FOLD_DISTINCT) v = (lw ^ rw) | (lw & rw & v);
    else if (p.fold == FOLD_NOT_DISTINCT) v = ~(lw | rw) | (lw & rw & v);
    if (p.fuse) {
      if (p.fold == FOLD_NONE) v &= lw & rw;
      unsigned c = __popcll(v);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(ACU_FULL_MASK, c, o);  // sum inside each 16-lane half = one 1024-row tile
      if ((lane & 15) == 0) p.tile_count[(sbase >> 10) + (lane >> 4)] = c;
    }
    p.out_bits[(sbase >> 6) + lane] = v;
    if (p.out_valid) {
      p.out_valid[(sbase >> 6) + lane] = lw & rw;
      valid_cnt += __popcll(lw & rw);
    }
  }
  if (p.out_valid) {
    valid_cnt = warp_sum(valid_cnt);
    if (lane == 0 && valid_cnt) atomicAdd(p.res + RES_COUNT, (unsigned long long)valid_cnt);
  }
}

// popcount of the 1024-row tiles [first_tile, n_tiles) of a plan mask (the ragged tail of a fused compare)
__global__ void __launch_bounds__(256) k_tile_counts(const uint64_t *__restrict__ mask, int64_t first_tile, int64_t n_tiles,
                                                     uint32_t *__restrict__ tile_count) {
  const int64_t t = first_tile + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  unsigned c = 0;
  for (int k = 0; k < 16; ++k) c += __popcll(mask[t * 16 + k]);
  tile_count[t] = c;
}

struct CmpFuse {  // destination of a fused compare -> filter plan
  uint64_t *mask;
  int64_t n_words_padded;
  uint32_t *tile_count;
  int64_t n_tiles;
};

template <class T>
acu_status cmp_typed(acu_ctx *ctx, acu_cmp_op op, const acu_array *l, const acu_array *r, acu_array_out *out, const CmpFuse *fuse = nullptr) {
  acu_array_out scratch_out{};
  if (fuse) out = &scratch_out;
  const bool ls = l->is_scalar != 0, rs = r->is_scalar != 0;
  if (l->len != r->len && !ls && !rs)  // cmp.rs:228-232
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0,
                    "Cannot compare arrays of different lengths, got %lld vs %lld", (long long)l->len, (long long)r->len);
  const int64_t len = ls ? r->len : l->len;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  if (len == 0) return ACU_OK;
  acu_status st;
  const int64_t lnc = acu_resolve_null_count(ctx, l, &st);
  ACU_TRY(st);
  const int64_t rnc = acu_resolve_null_count(ctx, r, &st);
  ACU_TRY(st);
  const bool ln = lnc > 0, rn = rnc > 0;  // logical_nulls().filter(null_count > 0)
  const bool fold = op == ACU_DISTINCT || op == ACU_NOT_DISTINCT;
  const bool l_null_scalar = ls && ln, r_null_scalar = rs && rn;
  if (fuse) {  // words past the data (the plan pads its mask to a multiple of 32 words) select nothing
    const int64_t used = (len + 63) / 64;
    if (fuse->n_words_padded > used)
      ACU_CUDA(ctx, cudaMemsetAsync(fuse->mask + used, 0, (size_t)(fuse->n_words_padded - used) * 8, ctx->stream));
  }
  if (!fold && (l_null_scalar || r_null_scalar) && !(ls && rs)) {
    // a null scalar against an array: BooleanArray::new_null(len) (cmp.rs:353, :364)
    if (fuse) {  // an all-null predicate selects nothing
      ACU_CUDA(ctx, cudaMemsetAsync(fuse->mask, 0, (size_t)fuse->n_words_padded * 8, ctx->stream));
      ACU_CUDA(ctx, cudaMemsetAsync(fuse->tile_count, 0, (size_t)fuse->n_tiles * 4, ctx->stream));
      return ACU_OK;
    }
    ACU_CUDA(ctx, cudaMemsetAsync(out->values, 0, acu_bitmap_bytes(len), ctx->stream));
    return set_new_null(ctx, len, 0, out);
  }
  CmpParams<T> p{};
  p.n = len;
  p.res = ctx->d_res;
  p.out_bits = fuse ? fuse->mask : static_cast<uint64_t *>(out->values);
  if (fuse) { p.fuse = 1; p.tile_count = fuse->tile_count; }
  // Less / Greater family: gt and lt_eq swap the operands (cmp.rs:481-488)
  const bool swap = (op == ACU_GT || op == ACU_LT_EQ);
  const acu_array *x = swap ? r : l, *y = swap ? l : r;
  const bool xs = x->is_scalar != 0, ys = y->is_scalar != 0;
  const bool xn = swap ? rn : ln, yn = swap ? ln : rn;
  p.a = static_cast<const T *>(x->values);
  p.b = static_cast<const T *>(y->values);
  p.a_scalar = xs && !(xs && ys);
  p.b_scalar = ys && !(xs && ys);
  p.neg = (op == ACU_NEQ || op == ACU_DISTINCT || op == ACU_LT_EQ || op == ACU_GT_EQ);
  p.fold = op == ACU_DISTINCT ? FOLD_DISTINCT : op == ACU_NOT_DISTINCT ? FOLD_NOT_DISTINCT : FOLD_NONE;
  if (xn) { if (xs && !(xs && ys)) p.a_null_scalar = 1; else { p.av = x->validity; p.aoff = x->validity_offset; } }
  if (yn) { if (ys && !(xs && ys)) p.b_null_scalar = 1; else { p.bv = y->validity; p.boff = y->validity_offset; } }
  if (!fold && (xn || yn) && !fuse) p.out_valid = reinterpret_cast<uint64_t *>(out->validity);
  int blk = 0;
  if (!fuse) {
    blk = acu_call_begin(ctx, &st);
    ACU_TRY(st);
    p.res = acu_dres(ctx, blk);
  }
  const bool lt = !(op == ACU_EQ || op == ACU_NEQ || fold);
  // streaming head over whole 2048-row super-groups when the value pointers are 16-B aligned
  const bool aligned = (p.a_scalar || (uintptr_t)p.a % 16 == 0) && (p.b_scalar || (uintptr_t)p.b % 16 == 0);
  const int64_t sgroups = aligned ? len / 2048 : 0;
  if (sgroups > 0) {
    const int64_t blocks = (sgroups + 7) / 8;
    if (lt) ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_cmp_v2<T, true>), acu_wave_grid(ctx, k_cmp_v2<T, true>, 256, 0, blocks), 256, 0, p, sgroups);
    else ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_cmp_v2<T, false>), acu_wave_grid(ctx, k_cmp_v2<T, false>, 256, 0, blocks), 256, 0, p, sgroups);
  }
  const int64_t head = sgroups * 2048;
  if (head < len) {  // ragged tail (or everything, for unaligned slices)
    CmpParams<T> q = p;
    q.n = len - head;
    if (!q.a_scalar) q.a += head;
    if (!q.b_scalar) q.b += head;
    q.aoff += head;
    q.boff += head;
    q.out_bits += head >> 6;
    if (q.out_valid) q.out_valid += head >> 6;
    const int64_t strips = (q.n + 63) >> 6;
    const int64_t blocks = (strips + 31) / 32;
    if (lt) ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_cmp<T, true>), acu_wave_grid(ctx, k_cmp<T, true>, 256, 0, blocks), 256, 0, q);
    else ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_cmp<T, false>), acu_wave_grid(ctx, k_cmp<T, false>, 256, 0, blocks), 256, 0, q);
    if (fuse) {  // the tail's tiles (head is a multiple of 2048 rows = 2 tiles)
      const int64_t first_tile = head >> 10, rest = fuse->n_tiles - first_tile;
      if (rest > 0) ACU_LAUNCH(ctx, k_tile_counts, (unsigned)((rest + 255) / 256), 256, 0, fuse->mask, first_tile, fuse->n_tiles, fuse->tile_count);
    }
  }
  if (fuse) return ACU_OK;  // stream-ordered: the plan's scan kernels follow on the same stream
  const bool has_valid = p.out_valid != nullptr;
  return acu_call_end(ctx, blk, [has_valid, len, out](const unsigned long long *h) -> acu_status {
    if (has_valid) {
      out->has_validity = 1;
      out->null_count = len - (int64_t)h[RES_COUNT];
    }
    return ACU_OK;
  });
}

// ---------------------------------------------------------------------------------------
// cast (numeric): unary_opt / try_unary over num_traits::cast (num-traits 0.2.19)
// ---------------------------------------------------------------------------------------
template <class I, class O>
__device__ __forceinline__ bool num_cast(I v, O &o) {
  if constexpr (is_fp<O>::value) {
    o = (O)v;  // cvt.rn: int->float RNE, f64->f32 RNE (overflow -> inf), always Some
    return true;
  } else if constexpr (is_fp<I>::value) {
    if (v != v) return false;
    constexpr bool OS = std::is_signed<O>::value;
    if constexpr (sizeof(I) > sizeof(O)) {
      const I lo = OS ? (I)std::numeric_limits<O>::min() - (I)1 : (I)-1;
      const I hi = (I)std::numeric_limits<O>::max() + (I)1;
      if (!(v > lo && v < hi)) return false;
    } else {
      const I hi = (I)std::numeric_limits<O>::max();
      if constexpr (OS) { if (!(v >= (I)std::numeric_limits<O>::min() && v < hi)) return false; }
      else { if (!(v > (I)-1 && v < hi)) return false; }
    }
    o = (O)v;  // truncates toward zero for in-range values
    return true;
  } else {
    constexpr bool IS = std::is_signed<I>::value, OS = std::is_signed<O>::value;
    if constexpr (IS == OS) {
      if (v < std::numeric_limits<O>::min() || v > std::numeric_limits<O>::max()) return false;
    } else if constexpr (IS) {
      if (v < 0) return false;
      if ((typename std::make_unsigned<I>::type)v > std::numeric_limits<O>::max()) return false;
    } else {
      if (v > (typename std::make_unsigned<O>::type)std::numeric_limits<O>::max()) return false;
    }
    o = (O)v;
    return true;
  }
}

template <class I, class O>
__global__ void __launch_bounds__(256) k_cast(const I *__restrict__ in, O *__restrict__ out, int64_t n,
                                              const uint8_t *__restrict__ iv, int64_t ioff, uint64_t *out_valid,
                                              int safe, unsigned long long *res) {
  constexpr int U = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t strips = (n + 63) >> 6;
  unsigned valid_cnt = 0;
  unsigned long long err = ~0ull;
  for (int64_t s0 = warp * U; s0 < strips; s0 += nwarps * U) {
    I v[U][2];
    uint64_t vw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = (s0 + u) * 64;
      vw[u] = ones_to(row, n);
      if (iv) vw[u] &= ld_bits64(iv, ioff + row, ioff + n);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t i = row + h * 32 + lane;
        v[u][h] = i < n ? __ldg(in + i) : I();
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = (s0 + u) * 64;
      if (row >= n) break;
      uint32_t failed[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t i = row + h * 32 + lane;
        const bool live = (vw[u] >> (h * 32 + lane)) & 1ull;  // unary_opt / try_unary: valid slots only
        O o = O();
        bool ok = true;
        if (live) {
          ok = num_cast<I, O>(v[u][h], o);
          if (!ok) { o = O(); if (!safe) { unsigned long long e = (unsigned long long)i; err = e < err ? e : err; } }
        }
        failed[h] = __ballot_sync(ACU_FULL_MASK, live && !ok);
        if (i < n) out[i] = o;
      }
      if (lane == 0 && out_valid) {
        uint64_t w = vw[u];
        if (safe) w &= ~((uint64_t)failed[0] | ((uint64_t)failed[1] << 32));  // unrepresentable => null
        out_valid[row >> 6] = w;
        valid_cnt += __popcll(w);
      }
    }
  }
  if (out_valid && lane == 0 && valid_cnt) atomicAdd(res + RES_COUNT, (unsigned long long)valid_cnt);
  if (err != ~0ull) atomicMin(res + RES_ERR_INDEX, err);
}

// Streaming variant for casts that cannot fail (anything -> float, widening integer casts):
// 2048-row super-groups, 128-bit input vectors, lane-owned validity words, zero under nulls.
template <class I, class O> struct cast_infallible {
  static constexpr bool value = is_fp<O>::value ||
      (!is_fp<I>::value && ((std::is_signed<I>::value == std::is_signed<O>::value && sizeof(O) >= sizeof(I)) ||
                            (!std::is_signed<I>::value && std::is_signed<O>::value && sizeof(O) > sizeof(I))));
};
template <class O, int EPL> struct alignas((sizeof(O) * EPL >= 16) ? 16 : sizeof(O) * EPL) OutPack { O v[EPL]; };

template <class I, class O>
__global__ void __launch_bounds__(256, 4) k_cast_v2(const I *__restrict__ in, O *__restrict__ out, const int64_t n,
                                                    const int64_t sgroups, const uint8_t *__restrict__ iv, const int64_t ioff,
                                                    uint64_t *__restrict__ out_valid, unsigned long long *__restrict__ res) {
  constexpr int EPL = 16 / sizeof(I);
  constexpr int RPL = 32 * EPL;
  constexpr int LOADS = 2048 / RPL;
  constexpr int U = LOADS < 4 ? LOADS : 4;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned valid_cnt = 0;
  for (int64_t sg = warp; sg < sgroups; sg += nwarps) {
    const int64_t sbase = sg << 11;
    uint64_t vw = ~0ull;
    if (iv) vw = ld_bits64(iv, ioff + sbase + lane * 64, ioff + n);
    if (out_valid) { out_valid[(sbase >> 6) + lane] = vw; valid_cnt += __popcll(vw); }
#pragma unroll 1
    for (int l0 = 0; l0 < LOADS; l0 += U) {
      Pack<I, EPL> v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = pack_load<I, EPL>(in + sbase + (int64_t)(l0 + u) * RPL + lane * EPL);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pos = (l0 + u) * RPL + lane * EPL;
        const uint64_t w = __shfl_sync(ACU_FULL_MASK, vw, pos >> 6);
        const uint32_t bits = (uint32_t)(w >> (pos & 63));
        OutPack<O, EPL> o;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          O x = O();
          if ((bits >> e) & 1u) num_cast<I, O>(v[u].v[e], x);  // unary_opt / try_unary: valid slots only, zero elsewhere
          o.v[e] = x;
        }
        *reinterpret_cast<OutPack<O, EPL> *>(out + sbase + pos) = o;
      }
    }
  }
  if (out_valid) {
    valid_cnt = warp_sum(valid_cnt);
    if (lane == 0 && valid_cnt) atomicAdd(res + RES_COUNT, (unsigned long long)valid_cnt);
  }
}

template <class I, class O>
acu_status cast_typed(acu_ctx *ctx, acu_dtype to, int32_t safe, const acu_array *a, acu_array_out *out) {
  const int64_t len = a->len;
  out->len = len;
  out->null_count = 0;
  // numeric_cast (safe) always carries a NullBuffer; try_numeric_cast clones the input's
  out->has_validity = safe ? 1 : (a->validity != nullptr);
  if (len == 0) return ACU_OK;
  uint64_t *ov = out->has_validity ? reinterpret_cast<uint64_t *>(out->validity) : nullptr;
  ACU_TRY(acu_res_reset(ctx));
  int64_t head = 0;
  if constexpr (cast_infallible<I, O>::value) {
    if ((uintptr_t)a->values % 16 == 0 && (uintptr_t)out->values % 16 == 0 && len >= 2048) {
      const int64_t sgroups = len / 2048;
      head = sgroups * 2048;
      ACU_LAUNCH_TIMED(ctx, ACU_K_CAST, (k_cast_v2<I, O>), acu_wave_grid(ctx, k_cast_v2<I, O>, 256, 0, (sgroups + 7) / 8), 256, 0,
                       static_cast<const I *>(a->values), static_cast<O *>(out->values), len, sgroups, a->validity,
                       a->validity_offset, ov, ctx->d_res);
    }
  }
  if (head < len) {
    const int64_t strips = (len - head + 63) >> 6;
    ACU_LAUNCH_TIMED(ctx, ACU_K_CAST, (k_cast<I, O>), acu_wave_grid(ctx, k_cast<I, O>, 256, 0, (strips + 31) / 32), 256, 0,
                     static_cast<const I *>(a->values) + head, static_cast<O *>(out->values) + head, len - head, a->validity,
                     a->validity_offset + head, ov ? ov + (head >> 6) : nullptr, safe, ctx->d_res);
  }
  ACU_TRY(acu_res_fetch(ctx));
  if (!safe && ctx->h_res[RES_ERR_INDEX] != ~0ull) {
    const int64_t idx = (int64_t)ctx->h_res[RES_ERR_INDEX];
    I v;
    ACU_CUDA(ctx, cudaMemcpyAsync(&v, static_cast<const I *>(a->values) + idx, sizeof(I), cudaMemcpyDeviceToHost, ctx->stream));
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    char s[40];
    fmt_native(s, sizeof s, v);
    return acu_fail(ctx, ACU_ERR_CAST, idx, bits_of(v), 0, 0, "Can't cast value %s to type %s", s, acu_dtype_name(to));
  }
  if (ov) out->null_count = len - (int64_t)ctx->h_res[RES_COUNT];
  return ACU_OK;
}

template <class I>
acu_status cast_from(acu_ctx *ctx, acu_dtype to, int32_t safe, const acu_array *a, acu_array_out *out) {
  switch (to) {
    case ACU_I8: return cast_typed<I, int8_t>(ctx, to, safe, a, out);
    case ACU_I16: return cast_typed<I, int16_t>(ctx, to, safe, a, out);
    case ACU_I32: return cast_typed<I, int32_t>(ctx, to, safe, a, out);
    case ACU_I64: return cast_typed<I, int64_t>(ctx, to, safe, a, out);
    case ACU_U8: return cast_typed<I, uint8_t>(ctx, to, safe, a, out);
    case ACU_U16: return cast_typed<I, uint16_t>(ctx, to, safe, a, out);
    case ACU_U32: return cast_typed<I, uint32_t>(ctx, to, safe, a, out);
    case ACU_U64: return cast_typed<I, uint64_t>(ctx, to, safe, a, out);
    case ACU_F32: return cast_typed<I, float>(ctx, to, safe, a, out);
    case ACU_F64: return cast_typed<I, double>(ctx, to, safe, a, out);
  }
  return acu_fail(ctx, ACU_ERR_NOT_YET_IMPLEMENTED, -1, 0, 0, 0, "cast to dtype %d", (int)to);
}

}  // namespace

#define ACU_DISPATCH(dt, F, ...)                         \
  switch (dt) {                                          \
    case ACU_I8: return F<int8_t>(__VA_ARGS__);          \
    case ACU_I16: return F<int16_t>(__VA_ARGS__);        \
    case ACU_I32: return F<int32_t>(__VA_ARGS__);        \
    case ACU_I64: return F<int64_t>(__VA_ARGS__);        \
    case ACU_U8: return F<uint8_t>(__VA_ARGS__);         \
    case ACU_U16: return F<uint16_t>(__VA_ARGS__);       \
    case ACU_U32: return F<uint32_t>(__VA_ARGS__);       \
    case ACU_U64: return F<uint64_t>(__VA_ARGS__);       \
    case ACU_F32: return F<float>(__VA_ARGS__);          \
    case ACU_F64: return F<double>(__VA_ARGS__);         \
  }

extern "C" acu_status acu_arith(acu_ctx *ctx, acu_dtype dtype, acu_arith_op op, const acu_array *a,
                                const acu_array *b, acu_array_out *out) {
  ACU_ENTER(ctx);
  ACU_DISPATCH(dtype, arith_typed, ctx, op, a, b, out)
  return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid arithmetic operation: dtype %d", (int)dtype);
}

extern "C" acu_status acu_neg(acu_ctx *ctx, acu_dtype dtype, int32_t checked, const acu_array *a, acu_array_out *out) {
  ACU_ENTER(ctx);
  if (checked && (dtype == ACU_U8 || dtype == ACU_U16 || dtype == ACU_U32 || dtype == ACU_U64))  // numeric.rs:174-176
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid arithmetic operation: !%s", acu_dtype_name(dtype));
  ACU_DISPATCH(dtype, neg_typed, ctx, checked, a, out)
  return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid arithmetic operation: dtype %d", (int)dtype);
}

extern "C" acu_status acu_cmp(acu_ctx *ctx, acu_dtype dtype, acu_cmp_op op, const acu_array *a,
                              const acu_array *b, acu_array_out *out) {
  ACU_ENTER(ctx);
  ACU_DISPATCH(dtype, cmp_typed, ctx, op, a, b, out)
  return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid comparison operation: dtype %d", (int)dtype);
}

// The comparison of acu_cmp written straight into a filter plan's mask / tile counts (compact.cu: acu_filter_plan_create_cmp).
// Stream-ordered, no synchronisation. *out_len = the result length (cmp.rs:228-235).
acu_status acu_cmp_result_len(acu_ctx *ctx, const acu_array *l, const acu_array *r, int64_t *out_len) {
  const bool ls = l->is_scalar != 0, rs = r->is_scalar != 0;
  if (l->len != r->len && !ls && !rs)
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0,
                    "Cannot compare arrays of different lengths, got %lld vs %lld", (long long)l->len, (long long)r->len);
  *out_len = ls ? r->len : l->len;
  return ACU_OK;
}
acu_status acu_cmp_into_plan(acu_ctx *ctx, acu_dtype dtype, acu_cmp_op op, const acu_array *a, const acu_array *b, uint64_t *mask,
                             int64_t n_words_padded, uint32_t *tile_count, int64_t n_tiles) {
  CmpFuse f{mask, n_words_padded, tile_count, n_tiles};
  const CmpFuse *fuse = &f;
  acu_array_out *out = nullptr;
  ACU_DISPATCH(dtype, cmp_typed, ctx, op, a, b, out, fuse)
  return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid comparison operation: dtype %d", (int)dtype);
}

extern "C" acu_status acu_cast_numeric(acu_ctx *ctx, acu_dtype from, acu_dtype to, int32_t safe,
                                       const acu_array *a, acu_array_out *out) {
  ACU_ENTER(ctx);
  ACU_DISPATCH(from, cast_from, ctx, to, safe, a, out)
  return acu_fail(ctx, ACU_ERR_NOT_YET_IMPLEMENTED, -1, 0, 0, 0, "cast from dtype %d", (int)from);
}
