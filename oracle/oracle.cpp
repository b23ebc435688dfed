// oracle.cpp — CPU restatement of the arrow-rs compute hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` leg may load this library. The product
// path (libarrow_cuda.so) never links, loads or calls anything in oracle/.
//
// Every function restates the algorithm of the cited reference file (paths relative to
// /root/reference, arrow-rs 59.2.0 @ cd7c6b83) over the exact arrow-buffer layout
// (values buffer + LSB-first validity bitmap + bit offset + cached null_count),
// single-threaded like the reference kernels, including what is written under null
// slots and when the result's NullBuffer is `None`.
//
// Parity pinning: the reference cannot be compiled here (no rustc). The oracle is pinned
// against the reference's own literal test vectors transcribed in tests/golden/*.json
// (see tests/test_oracle_golden.py), not against outputs of the reference itself.
// Third-party arithmetic: num-traits 0.2.19 `cast::<I,O>` (Cargo.lock) is restated in
// num_cast<>() below from its published semantics and pinned by
// arrow-cast/src/cast/mod.rs:8449-8481 (test_cast_from_int64).
//
// Build: see oracle/Makefile  (g++ -O3 -march=x86-64-v3 -ffp-contract=off: Rust never
// contracts a*b+c, and neither may we).

#include <cinttypes>
#include <cstdarg>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

#include "../include/arrow_cuda.h"

namespace {

thread_local acu_error_detail g_err;

acu_status fail(acu_status st, int64_t index, uint64_t lhs, uint64_t rhs, uint64_t len,
                const char *fmt, ...) __attribute__((format(printf, 6, 7)));
acu_status fail(acu_status st, int64_t index, uint64_t lhs, uint64_t rhs, uint64_t len,
                const char *fmt, ...) {
  g_err.status = st;
  g_err.cuda_error = 0;
  g_err.index = index;
  g_err.lhs_bits = lhs;
  g_err.rhs_bits = rhs;
  g_err.len = len;
  // message = Display of the ArrowError variant (arrow-schema/src/error.rs:96-137)
  const char *prefix = "";
  switch (st) {
    case ACU_ERR_INVALID_ARGUMENT: prefix = "Invalid argument error: "; break;
    case ACU_ERR_COMPUTE: prefix = "Compute error: "; break;
    case ACU_ERR_ARITHMETIC_OVERFLOW: prefix = "Arithmetic overflow: "; break;
    case ACU_ERR_OFFSET_OVERFLOW: prefix = "Offset overflow error: "; break;
    case ACU_ERR_CAST: prefix = "Cast error: "; break;
    default: break;
  }
  size_t n = strlen(prefix);
  memcpy(g_err.message, prefix, n);
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err.message + n, sizeof(g_err.message) - n, fmt, ap);
  va_end(ap);
  return st;
}

// ---- bit utilities: arrow-buffer/src/util/bit_util.rs:52-115 -------------------------
inline bool get_bit(const uint8_t *b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
inline void set_bit(uint8_t *b, int64_t i) { b[i >> 3] |= (uint8_t)(1u << (i & 7)); }

// 64 bits starting at bit `pos` (bits past `end` read as 0). Restates the padded
// chunk iteration of BitChunks::iter_padded (arrow-buffer/src/util/bit_chunk_iterator.rs:220-337).
inline uint64_t load_bits(const uint8_t *b, int64_t pos, int64_t end) {
  uint64_t w = 0;
  int64_t n = end - pos;
  if (n <= 0) return 0;
  if (n > 64) n = 64;
  int64_t byte0 = pos >> 3;
  int shift = (int)(pos & 7);
  int64_t nbytes = (shift + n + 7) >> 3;  // <= 9
  uint64_t lo = 0;
  uint8_t hi = 0;
  for (int64_t k = 0; k < nbytes && k < 8; ++k) lo |= (uint64_t)b[byte0 + k] << (8 * k);
  if (nbytes == 9) hi = b[byte0 + 8];
  w = shift ? (lo >> shift) | ((uint64_t)hi << (64 - shift)) : lo;
  if (n < 64) w &= (~0ull) >> (64 - n);
  return w;
}

inline int64_t count_bits(const uint8_t *b, int64_t off, int64_t len) {
  int64_t c = 0;
  for (int64_t i = 0; i < len; i += 64) c += __builtin_popcountll(load_bits(b, off + i, off + len));
  return c;
}

inline int64_t resolve_null_count(const acu_array *a) {
  if (!a->validity) return 0;
  if (a->null_count >= 0) return a->null_count;
  return a->len - count_bits(a->validity, a->validity_offset, a->len);
}

// ---- filter: arrow-select/src/filter.rs ----------------------------------------------

constexpr double FILTER_SLICES_SELECTIVITY_THRESHOLD = 0.8;  // filter.rs:43

struct Predicate {
  std::vector<uint64_t> mask;  // normalised to bit offset 0 (prep_null_mask_filter :167-171)
  int64_t len = 0;
  int64_t count = 0;
  int32_t strategy = ACU_FILTER_NONE;
  const uint8_t *bytes() const { return reinterpret_cast<const uint8_t *>(mask.data()); }
};

// FilterBuilder::new (filter.rs:254-273): pass #1 true_count
// (arrow-array/src/array/boolean_array.rs:175-187), then `values & nulls` when the
// predicate has nulls, then IterationStrategy::default_strategy (filter.rs:346-364).
void build_predicate(const acu_array *p, Predicate *out) {
  const uint8_t *vals = static_cast<const uint8_t *>(p->values);
  int64_t len = p->len;
  out->len = len;
  out->mask.assign((size_t)((len + 63) / 64) + 1, 0);
  bool has_nulls = p->validity && resolve_null_count(p) > 0;
  int64_t count = 0;
  for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
    uint64_t v = load_bits(vals, p->values_offset + i, p->values_offset + len);
    if (p->validity)  // true_count ANDs the validity whenever a NullBuffer exists
      count += __builtin_popcountll(
          v & load_bits(p->validity, p->validity_offset + i, p->validity_offset + len));
    else
      count += __builtin_popcountll(v);
    if (has_nulls) v &= load_bits(p->validity, p->validity_offset + i, p->validity_offset + len);
    out->mask[w] = v;
  }
  out->count = count;
  if (len == 0 || count == 0)
    out->strategy = ACU_FILTER_NONE;
  else if (count == len)
    out->strategy = ACU_FILTER_ALL;
  else if ((double)count / (double)len > FILTER_SLICES_SELECTIVITY_THRESHOLD)
    out->strategy = ACU_FILTER_SLICES;
  else
    out->strategy = ACU_FILTER_INDEX;
}

// BitIndexIterator (arrow-buffer/src/util/bit_iterator.rs:284-325): trailing_zeros +
// clear-lowest-set-bit over u64 chunks.
template <class F>
inline void for_each_index(const Predicate &p, F &&f) {
  size_t words = (size_t)((p.len + 63) / 64);
  for (size_t w = 0; w < words; ++w) {
    uint64_t m = p.mask[w];
    while (m) {
      int tz = __builtin_ctzll(m);
      f((int64_t)w * 64 + tz);
      m &= m - 1;
    }
  }
}

// SlicesIterator / BitSliceIterator (bit_iterator.rs:186-283): maximal runs of set bits.
template <class F>
inline void for_each_slice(const Predicate &p, F &&f) {
  int64_t i = 0, len = p.len;
  const uint8_t *b = p.bytes();
  while (i < len) {
    // skip zeros
    uint64_t w = load_bits(b, i, len);
    if (w == 0) { i += 64; continue; }
    int64_t start = i + __builtin_ctzll(w);
    int64_t j = start;
    for (;;) {
      uint64_t x = ~load_bits(b, j, len);  // zeros (and past-the-end) become ones
      if (len - j < 64) x |= (~0ull) << (len - j);
      if (x == 0) { j += 64; continue; }
      j += __builtin_ctzll(x);
      break;
    }
    f(start, j);
    i = j;
  }
}

// filter_bits (filter.rs:680-720): compact the bits of `src` (bit offset `off`) selected
// by the predicate into a fresh bitmap with bit offset 0.
void filter_bits(const Predicate &p, const uint8_t *src, int64_t off, uint8_t *dst) {
  memset(dst, 0, acu_bitmap_bytes(p.count));
  int64_t k = 0;
  if (p.strategy == ACU_FILTER_INDEX) {
    for_each_index(p, [&](int64_t i) {
      if (get_bit(src, off + i)) set_bit(dst, k);
      ++k;
    });
  } else {  // append_packed_range per slice
    for_each_slice(p, [&](int64_t s, int64_t e) {
      for (int64_t i = s; i < e; ++i, ++k)
        if (get_bit(src, off + i)) set_bit(dst, k);
    });
  }
}

// FilterPredicate::filter_nulls (filter.rs:512-533)
void filter_nulls(const Predicate &p, const acu_array *a, acu_array_out *out) {
  out->has_validity = 0;
  out->null_count = 0;
  if (!a->validity) return;
  if (resolve_null_count(a) == 0) return;
  filter_bits(p, a->validity, a->validity_offset, out->validity);
  int64_t nc = p.count - count_bits(out->validity, 0, p.count);
  if (nc == 0) return;
  out->has_validity = 1;
  out->null_count = nc;
}

// `values.slice(0, count)` for IterationStrategy::All (filter.rs:546): the reference
// shares the buffers; the oracle materialises the same logical array.
void slice_nulls(const acu_array *a, int64_t count, acu_array_out *out) {
  out->has_validity = 0;
  out->null_count = 0;
  if (!a->validity) return;
  memset(out->validity, 0, acu_bitmap_bytes(count));
  for (int64_t i = 0; i < count; ++i)
    if (get_bit(a->validity, a->validity_offset + i)) set_bit(out->validity, i);
  out->has_validity = 1;  // NullBuffer::slice keeps Some(..) with a recomputed null_count
  out->null_count = count - count_bits(out->validity, 0, count);
}

acu_status check_filter_len(const Predicate &p, int64_t values_len) {
  if (p.len > values_len)  // filter.rs:536-542
    return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, (uint64_t)values_len,
                "Filter predicate of length %" PRId64
                " is larger than target array of length %" PRId64,
                p.len, values_len);
  return ACU_OK;
}

// ---- take: arrow-select/src/take.rs -----------------------------------------------------

// ToIndices (take.rs:1030-1084): i8/i16/u8/u16 widen with `as u32` (sign-extending),
// i32 -> u32 and i64 -> u64 reinterpret. Returned as u64 "as_usize".
inline uint64_t load_index(const void *idx, acu_dtype t, int64_t j) {
  switch (t) {
    case ACU_I8: return (uint32_t)(int32_t) static_cast<const int8_t *>(idx)[j];
    case ACU_I16: return (uint32_t)(int32_t) static_cast<const int16_t *>(idx)[j];
    case ACU_U8: return static_cast<const uint8_t *>(idx)[j];
    case ACU_U16: return static_cast<const uint16_t *>(idx)[j];
    case ACU_I32:
    case ACU_U32: return static_cast<const uint32_t *>(idx)[j];
    case ACU_I64:
    case ACU_U64: return static_cast<const uint64_t *>(idx)[j];
    default: return ~0ull;
  }
}
inline bool index_is_signed(acu_dtype t) { return t == ACU_I8 || t == ACU_I16 || t == ACU_I32 || t == ACU_I64; }
inline int64_t load_index_signed(const void *idx, acu_dtype t, int64_t j) {
  switch (t) {
    case ACU_I8: return static_cast<const int8_t *>(idx)[j];
    case ACU_I16: return static_cast<const int16_t *>(idx)[j];
    case ACU_I32: return static_cast<const int32_t *>(idx)[j];
    case ACU_I64: return static_cast<const int64_t *>(idx)[j];
    default: return 0;
  }
}
inline uint64_t dtype_max(acu_dtype t) {
  switch (t) {
    case ACU_I8: return INT8_MAX; case ACU_I16: return INT16_MAX; case ACU_I32: return INT32_MAX;
    case ACU_I64: return INT64_MAX; case ACU_U8: return UINT8_MAX; case ACU_U16: return UINT16_MAX;
    case ACU_U32: return UINT32_MAX; default: return UINT64_MAX;
  }
}
inline bool is_int_dtype(acu_dtype t) { return t <= ACU_U64; }

// check_bounds (take.rs:167-209): on the ORIGINAL index type, before ToIndices.
acu_status check_bounds(int64_t len, const acu_array *ind, acu_dtype t) {
  if ((uint64_t)len > dtype_max(t)) return ACU_OK;  // T::Native::from_usize(len) is None
  bool has_nulls = ind->validity && resolve_null_count(ind) > 0;
  for (int64_t j = 0; j < ind->len; ++j) {
    if (has_nulls && !get_bit(ind->validity, ind->validity_offset + j)) continue;
    bool bad;
    char buf[32];
    if (index_is_signed(t)) {
      int64_t v = load_index_signed(ind->values, t, j);
      // nullable path tests only `index >= len` (take.rs:183); the non-null path also
      // tests `index < 0` (take.rs:193-199)
      bad = has_nulls ? (v >= len) : (v < 0 || v >= len);
      snprintf(buf, sizeof buf, "%" PRId64, v);
    } else {
      uint64_t v = load_index(ind->values, t, j);
      bad = v >= (uint64_t)len;
      snprintf(buf, sizeof buf, "%" PRIu64, v);
    }
    if (bad)
      return fail(ACU_ERR_COMPUTE, j, load_index(ind->values, t, j), 0, (uint64_t)len,
                  "Array index out of bounds, cannot get item at index %s from %" PRId64
                  " entries",
                  buf, len);
  }
  return ACU_OK;
}

// take_nulls (take.rs:419-430) + take_bits (take.rs:460-486)
void take_nulls(const acu_array *values, const acu_array *ind, acu_dtype t, acu_array_out *out) {
  int64_t m = ind->len;
  bool ind_nulls = ind->validity && resolve_null_count(ind) > 0;
  out->has_validity = 0;
  out->null_count = 0;
  if (values->validity && resolve_null_count(values) > 0) {
    memset(out->validity, 0, acu_bitmap_bytes(m));
    for (int64_t j = 0; j < m; ++j) {
      if (ind_nulls && !get_bit(ind->validity, ind->validity_offset + j)) continue;
      uint64_t ix = load_index(ind->values, t, j);
      if (get_bit(values->validity, values->validity_offset + (int64_t)ix)) set_bit(out->validity, j);
    }
    int64_t nc = m - count_bits(out->validity, 0, m);
    if (nc > 0) {  // NullBuffer::from_unsliced_buffer (null.rs:266-270)
      out->has_validity = 1;
      out->null_count = nc;
    }
  } else if (ind->validity) {  // indices.nulls().cloned()
    memset(out->validity, 0, acu_bitmap_bytes(m));
    for (int64_t j = 0; j < m; ++j)
      if (get_bit(ind->validity, ind->validity_offset + j)) set_bit(out->validity, j);
    out->has_validity = 1;
    out->null_count = m - count_bits(out->validity, 0, m);
  }
}

// Validates what the reference would panic on (take.rs:447, :454 and the bounds-checked
// BooleanBuffer::value in take_bits): an out-of-bounds index in a VALID slot.
acu_status check_panic(int64_t values_len, const acu_array *ind, acu_dtype t) {
  bool ind_nulls = ind->validity && resolve_null_count(ind) > 0;
  for (int64_t j = 0; j < ind->len; ++j) {
    uint64_t ix = load_index(ind->values, t, j);
    if (ix < (uint64_t)values_len) continue;
    if (ind_nulls && !get_bit(ind->validity, ind->validity_offset + j)) continue;
    return fail(ACU_ERR_PANIC_OUT_OF_BOUNDS, j, ix, 0, (uint64_t)values_len,
                "Out-of-bounds index %" PRIu64, ix);
  }
  return ACU_OK;
}

// ---- native type semantics: arrow-array/src/arithmetic.rs -----------------------------

template <class T> struct Bits;  // bit pattern helpers
template <class T> inline uint64_t to_bits(T v) {
  if constexpr (sizeof(T) == 8) { uint64_t u; memcpy(&u, &v, 8); return u; }
  else if constexpr (sizeof(T) == 4) { uint32_t u; memcpy(&u, &v, 4); return u; }
  else if constexpr (sizeof(T) == 2) { uint16_t u; memcpy(&u, &v, 2); return u; }
  else { uint8_t u; memcpy(&u, &v, 1); return u; }
}
template <class T> inline T from_bits(uint64_t b) {
  T v;
  if constexpr (sizeof(T) == 8) { memcpy(&v, &b, 8); }
  else if constexpr (sizeof(T) == 4) { uint32_t u = (uint32_t)b; memcpy(&v, &u, 4); }
  else if constexpr (sizeof(T) == 2) { uint16_t u = (uint16_t)b; memcpy(&v, &u, 2); }
  else { uint8_t u = (uint8_t)b; memcpy(&v, &u, 1); }
  return v;
}

// f64::total_cmp / f32::total_cmp key (arithmetic.rs:400-410, Rust core)
inline int64_t total_key(double x) {
  int64_t b = (int64_t)to_bits(x);
  return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}
inline int32_t total_key(float x) {
  int32_t b = (int32_t)to_bits(x);
  return b ^ (int32_t)((uint32_t)(b >> 31) >> 1);
}
template <class T> inline bool is_eq(T l, T r) {
  if constexpr (std::is_floating_point<T>::value) return to_bits(l) == to_bits(r);
  else return l == r;
}
template <class T> inline bool is_lt(T l, T r) {
  if constexpr (std::is_floating_point<T>::value) return total_key(l) < total_key(r);
  else return l < r;
}
template <class T> inline bool is_gt(T l, T r) { return is_lt(r, l); }

template <class T> void fmt_native(char *buf, size_t n, T v) {  // Rust {:?}
  if constexpr (std::is_floating_point<T>::value) snprintf(buf, n, "%.17g", (double)v);
  else if constexpr (std::is_signed<T>::value) snprintf(buf, n, "%" PRId64, (int64_t)v);
  else snprintf(buf, n, "%" PRIu64, (uint64_t)v);
}

enum class OpErr { None, Overflow, DivZero };

// One element of integer_op / float_op (numeric.rs:328-374).
template <class T>
inline OpErr apply_arith(acu_arith_op op, T l, T r, T *out) {
  if constexpr (std::is_floating_point<T>::value) {
    switch (op) {
      case ACU_ADD_WRAPPING: case ACU_ADD: *out = l + r; break;
      case ACU_SUB_WRAPPING: case ACU_SUB: *out = l - r; break;
      case ACU_MUL_WRAPPING: case ACU_MUL: *out = l * r; break;
      case ACU_DIV: *out = l / r; break;
      case ACU_REM: *out = std::fmod(l, r); break;  // Rust `%` on floats
    }
    return OpErr::None;
  } else {
    using U = typename std::make_unsigned<T>::type;
    switch (op) {
      case ACU_ADD_WRAPPING: *out = (T)((U)l + (U)r); return OpErr::None;
      case ACU_SUB_WRAPPING: *out = (T)((U)l - (U)r); return OpErr::None;
      case ACU_MUL_WRAPPING: *out = (T)((U)l * (U)r); return OpErr::None;
      case ACU_ADD: return __builtin_add_overflow(l, r, out) ? OpErr::Overflow : OpErr::None;
      case ACU_SUB: return __builtin_sub_overflow(l, r, out) ? OpErr::Overflow : OpErr::None;
      case ACU_MUL: return __builtin_mul_overflow(l, r, out) ? OpErr::Overflow : OpErr::None;
      case ACU_DIV:  // div_checked (arithmetic.rs:204-215)
        if (r == 0) return OpErr::DivZero;
        if (std::is_signed<T>::value && l == std::numeric_limits<T>::min() && r == (T)-1)
          return OpErr::Overflow;
        *out = (T)(l / r);
        return OpErr::None;
      case ACU_REM:  // numeric.rs:345-351: zero => DivideByZero else wrapping_rem
        if (r == 0) return OpErr::DivZero;
        if (std::is_signed<T>::value && r == (T)-1) { *out = 0; return OpErr::None; }
        *out = (T)(l % r);
        return OpErr::None;
    }
    return OpErr::None;
  }
}

inline const char *op_symbol(acu_arith_op op) {  // numeric.rs:192-202
  switch (op) {
    case ACU_ADD_WRAPPING: case ACU_ADD: return "+";
    case ACU_SUB_WRAPPING: case ACU_SUB: return "-";
    case ACU_MUL_WRAPPING: case ACU_MUL: return "*";
    case ACU_DIV: return "/";
    default: return "%";
  }
}

template <class T>
acu_status arith_error(OpErr e, acu_arith_op op, int64_t idx, T l, T r) {
  if (e == OpErr::DivZero)  // arrow-schema/src/error.rs Display for DivideByZero
    return fail(ACU_ERR_DIVIDE_BY_ZERO, idx, to_bits(l), to_bits(r), 0, "Divide by zero error");
  char a[40], b[40];
  fmt_native(a, sizeof a, l);
  fmt_native(b, sizeof b, r);
  return fail(ACU_ERR_ARITHMETIC_OVERFLOW, idx, to_bits(l), to_bits(r), 0,
              "Overflow happened on: %s %s %s", a, op_symbol(op), b);
}

inline bool op_is_checked_int(acu_arith_op op) {
  return op == ACU_ADD || op == ACU_SUB || op == ACU_MUL || op == ACU_DIV || op == ACU_REM;
}

// NullBuffer::union (arrow-buffer/src/buffer/null.rs:79-87). Returns true if Some.
bool null_union(const acu_array *a, const acu_array *b, int64_t len, acu_array_out *out) {
  int64_t an = resolve_null_count(a), bn = resolve_null_count(b);
  out->has_validity = 0;
  out->null_count = 0;
  const acu_array *one = nullptr;
  if (a->validity && b->validity) {
    if (an == 0 && bn == 0) return false;
    for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
      uint64_t x = load_bits(a->validity, a->validity_offset + i, a->validity_offset + len) &
                   load_bits(b->validity, b->validity_offset + i, b->validity_offset + len);
      memcpy(out->validity + 8 * w, &x, 8);
    }
  } else if (a->validity && an > 0) {
    one = a;
  } else if (b->validity && bn > 0) {
    one = b;
  } else {
    return false;
  }
  if (one) {
    for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
      uint64_t x = load_bits(one->validity, one->validity_offset + i, one->validity_offset + len);
      memcpy(out->validity + 8 * w, &x, 8);
    }
  }
  out->has_validity = 1;
  out->null_count = len - count_bits(out->validity, 0, len);
  return true;
}

void clone_nulls(const acu_array *a, int64_t len, acu_array_out *out) {  // nulls().cloned()
  out->has_validity = 0;
  out->null_count = 0;
  if (!a->validity) return;
  for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
    uint64_t x = load_bits(a->validity, a->validity_offset + i, a->validity_offset + len);
    memcpy(out->validity + 8 * w, &x, 8);
  }
  out->has_validity = 1;
  out->null_count = len - count_bits(out->validity, 0, len);
}

// PrimitiveArray::new_null(len) (primitive_array.rs:658-664)
template <class T> void new_null(int64_t len, acu_array_out *out) {
  memset(out->values, 0, (size_t)len * sizeof(T));
  memset(out->validity, 0, acu_bitmap_bytes(len));
  out->len = len;
  out->has_validity = 1;
  out->null_count = len;
}

template <class T>
acu_status arith_typed(acu_arith_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  const T *av = static_cast<const T *>(a->values);
  const T *bv = static_cast<const T *>(b->values);
  T *ov = static_cast<T *>(out->values);
  const bool checked = !std::is_floating_point<T>::value && op_is_checked_int(op);
  const bool a_s = a->is_scalar != 0, b_s = b->is_scalar != 0;

  if (a_s != b_s) {  // op!/try_op! scalar arms (numeric.rs:278-317)
    const acu_array *s = a_s ? a : b;
    const acu_array *arr = a_s ? b : a;
    int64_t len = arr->len;
    out->len = len;
    if (resolve_null_count(s) != 0) { new_null<T>(len, out); return ACU_OK; }
    T sv = static_cast<const T *>(s->values)[0];
    const T *xv = static_cast<const T *>(arr->values);
    clone_nulls(arr, len, out);
    if (!checked) {  // PrimitiveArray::unary (primitive_array.rs:916-925): every slot
      for (int64_t i = 0; i < len; ++i) {
        T l = a_s ? sv : xv[i], r = a_s ? xv[i] : sv;
        apply_arith(op, l, r, &ov[i]);
      }
      return ACU_OK;
    }
    // try_unary (primitive_array.rs:990-1014): zero-filled, valid slots only
    memset(ov, 0, (size_t)len * sizeof(T));
    if (arr->validity && out->null_count == len) return ACU_OK;
    for (int64_t i = 0; i < len; ++i) {
      if (arr->validity && !get_bit(arr->validity, arr->validity_offset + i)) continue;
      T l = a_s ? sv : xv[i], r = a_s ? xv[i] : sv;
      OpErr e = apply_arith(op, l, r, &ov[i]);
      if (e != OpErr::None) return arith_error(e, op, i, l, r);
    }
    return ACU_OK;
  }

  // array ∘ array (or scalar ∘ scalar): binary / try_binary (arity.rs:104-135, :254-299)
  if (a->len != b->len)
    return fail(ACU_ERR_COMPUTE, -1, 0, 0, 0,
                checked ? "Cannot perform a binary operation on arrays of different length"
                        : "Cannot perform binary operation on arrays of different length");
  int64_t len = a->len;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  if (len == 0) return ACU_OK;
  if (!checked) {
    null_union(a, b, len, out);
    for (int64_t i = 0; i < len; ++i) apply_arith(op, av[i], bv[i], &ov[i]);
    return ACU_OK;
  }
  bool has_union = false;
  if (resolve_null_count(a) != 0 || resolve_null_count(b) != 0) has_union = null_union(a, b, len, out);
  if (!has_union) {  // try_binary_no_nulls (arity.rs:383-400)
    for (int64_t i = 0; i < len; ++i) {
      OpErr e = apply_arith(op, av[i], bv[i], &ov[i]);
      if (e != OpErr::None) return arith_error(e, op, i, av[i], bv[i]);
    }
    return ACU_OK;
  }
  memset(ov, 0, (size_t)len * sizeof(T));
  if (out->null_count == len) return ACU_OK;  // try_for_each_valid_idx (null.rs:235-243)
  for (int64_t i = 0; i < len; ++i) {
    if (!get_bit(out->validity, i)) continue;
    OpErr e = apply_arith(op, av[i], bv[i], &ov[i]);
    if (e != OpErr::None) return arith_error(e, op, i, av[i], bv[i]);
  }
  return ACU_OK;
}

template <class T>
acu_status neg_typed(int32_t checked, const acu_array *a, acu_array_out *out) {
  const T *av = static_cast<const T *>(a->values);
  T *ov = static_cast<T *>(out->values);
  int64_t len = a->len;
  out->len = len;
  clone_nulls(a, len, out);
  if constexpr (std::is_floating_point<T>::value) {
    for (int64_t i = 0; i < len; ++i) ov[i] = -av[i];  // neg_wrapping! via unary
    return ACU_OK;
  } else {
    using U = typename std::make_unsigned<T>::type;
    if (!checked) {  // neg_wrapping: unary(wrapping_neg)
      for (int64_t i = 0; i < len; ++i) ov[i] = (T)((U)0 - (U)av[i]);
      return ACU_OK;
    }
    memset(ov, 0, (size_t)len * sizeof(T));  // neg_checked! via try_unary
    if (a->validity && out->null_count == len) return ACU_OK;
    for (int64_t i = 0; i < len; ++i) {
      if (a->validity && !get_bit(a->validity, a->validity_offset + i)) continue;
      if (av[i] == std::numeric_limits<T>::min()) {
        char s[40];
        fmt_native(s, sizeof s, av[i]);
        return fail(ACU_ERR_ARITHMETIC_OVERFLOW, i, to_bits(av[i]), 0, 0,
                    "Overflow happened on: - %s", s);
      }
      ov[i] = (T)(-av[i]);
    }
    return ACU_OK;
  }
}

// ---- cmp: arrow-ord/src/cmp.rs -----------------------------------------------------------

// collect_bool (cmp.rs:580-611): 64 results per u64, LSB first, whole-word negation
// (bits past len are left set when neg — "unspecified" for callers).
template <class F>
void collect_bool(int64_t len, bool neg, uint8_t *out, F &&f) {
  int64_t chunks = len / 64, rem = len % 64;
  for (int64_t c = 0; c < chunks; ++c) {
    uint64_t packed = 0;
    for (int bit = 0; bit < 64; ++bit) packed |= (uint64_t)f(c * 64 + bit) << bit;
    if (neg) packed = ~packed;
    memcpy(out + 8 * c, &packed, 8);
  }
  if (rem) {
    uint64_t packed = 0;
    for (int bit = 0; bit < rem; ++bit) packed |= (uint64_t)f(chunks * 64 + bit) << bit;
    if (neg) packed = ~packed;
    memcpy(out + 8 * chunks, &packed, 8);
  }
}

template <class T>
void cmp_values(acu_cmp_op op, const acu_array *l, const acu_array *r, int64_t len, uint8_t *out) {
  const T *lv = static_cast<const T *>(l->values);
  const T *rv = static_cast<const T *>(r->values);
  const bool ls = l->is_scalar != 0, rs = r->is_scalar != 0;
  auto L = [&](int64_t i) { return ls ? lv[0] : lv[i]; };
  auto R = [&](int64_t i) { return rs ? rv[0] : rv[i]; };
  switch (op) {  // apply (cmp.rs:481-488)
    case ACU_EQ: case ACU_NOT_DISTINCT: collect_bool(len, false, out, [&](int64_t i) { return is_eq(L(i), R(i)); }); break;
    case ACU_NEQ: case ACU_DISTINCT: collect_bool(len, true, out, [&](int64_t i) { return is_eq(L(i), R(i)); }); break;
    case ACU_LT: collect_bool(len, false, out, [&](int64_t i) { return is_lt(L(i), R(i)); }); break;
    case ACU_LT_EQ: collect_bool(len, true, out, [&](int64_t i) { return is_lt(R(i), L(i)); }); break;
    case ACU_GT: collect_bool(len, false, out, [&](int64_t i) { return is_lt(R(i), L(i)); }); break;
    case ACU_GT_EQ: collect_bool(len, true, out, [&](int64_t i) { return is_lt(L(i), R(i)); }); break;
  }
}

inline uint64_t side_bits(const acu_array *s, int64_t i, int64_t len) {  // bit_chunks().iter_padded()
  if (s->is_scalar) return get_bit(s->validity, s->validity_offset) ? 1ull : 0ull;
  return load_bits(s->validity, s->validity_offset + i, s->validity_offset + len);
}

// compare_op (cmp.rs:220-382) for any operand kind: `l` / `r` carry len / validity / scalar-ness, `cmp_vals(len, out)` fills
// the value bits through apply() (cmp.rs:438-502).
template <class V>
acu_status cmp_generic(acu_cmp_op op, const acu_array *l, const acu_array *r, acu_array_out *out, V &&cmp_vals) {
  const bool ls = l->is_scalar != 0, rs = r->is_scalar != 0;
  if (l->len != r->len && !ls && !rs)  // cmp.rs:228-232
    return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0,
                "Cannot compare arrays of different lengths, got %" PRId64 " vs %" PRId64, l->len, r->len);
  int64_t len = ls ? r->len : l->len;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  uint8_t *ov = static_cast<uint8_t *>(out->values);
  const bool ln = l->validity && resolve_null_count(l) > 0;
  const bool rn = r->validity && resolve_null_count(r) > 0;
  const size_t obytes = acu_bitmap_bytes(len);
  auto values = [&]() {
    if (l->len == 0 || r->len == 0) memset(ov, 0, obytes);  // apply returns None => new_unset
    else cmp_vals(len, ov);
  };
  auto new_null_bool = [&]() {  // BooleanArray::new_null(len)
    memset(ov, 0, obytes);
    memset(out->validity, 0, obytes);
    out->has_validity = 1;
    out->null_count = len;
  };
  if (ln && rn && ls == rs) {  // cmp.rs:320-346
    if (op == ACU_DISTINCT || op == ACU_NOT_DISTINCT) {
      values();
      for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
        uint64_t a = side_bits(l, i, len), b = side_bits(r, i, len), v;
        memcpy(&v, ov + 8 * w, 8);
        if (len - i < 64) v &= (~0ull) >> (64 - (len - i));  // iter_padded zero-pads
        uint64_t c = op == ACU_DISTINCT ? ((a ^ b) | (a & b & v)) : (~(a | b) | (a & b & v));
        memcpy(ov + 8 * w, &c, 8);
      }
    } else {
      values();
      // both sides nullable and same scalar-ness: union = AND of the two bitmaps
      for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
        uint64_t x = side_bits(l, i, len) & side_bits(r, i, len);
        memcpy(out->validity + 8 * w, &x, 8);
      }
      out->has_validity = 1;
      out->null_count = len - count_bits(out->validity, 0, len);
    }
  } else if (ln && rn) {  // one side is a null scalar, the other nullable (cmp.rs:348-355)
    const acu_array *a = ls ? r : l;
    if (op == ACU_DISTINCT || op == ACU_NOT_DISTINCT) {
      for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
        uint64_t x = load_bits(a->validity, a->validity_offset + i, a->validity_offset + len);
        if (op == ACU_NOT_DISTINCT) x = ~x;
        memcpy(ov + 8 * w, &x, 8);
      }
    } else {
      new_null_bool();
    }
  } else if (ln || rn) {  // only one side nullable (cmp.rs:356-378)
    const acu_array *n = ln ? l : r;
    if (n->is_scalar) {
      if (op == ACU_DISTINCT) memset(ov, 0xff, obytes);
      else if (op == ACU_NOT_DISTINCT) memset(ov, 0, obytes);
      else new_null_bool();
    } else if (op == ACU_DISTINCT || op == ACU_NOT_DISTINCT) {
      values();
      for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
        uint64_t a = load_bits(n->validity, n->validity_offset + i, n->validity_offset + len), v;
        memcpy(&v, ov + 8 * w, 8);
        if (len - i < 64) v &= (~0ull) >> (64 - (len - i));
        uint64_t c = op == ACU_DISTINCT ? (~a | v) : (a & v);
        memcpy(ov + 8 * w, &c, 8);
      }
    } else {
      values();
      clone_nulls(n, len, out);
    }
  } else {
    values();
  }
  return ACU_OK;
}

template <class T>
acu_status cmp_typed(acu_cmp_op op, const acu_array *l, const acu_array *r, acu_array_out *out) {
  return cmp_generic(op, l, r, out, [&](int64_t len, uint8_t *ov) { cmp_values<T>(op, l, r, len, ov); });
}

// apply() over any item accessor: eq / lt closures on logical row indices (cmp.rs:481-488)
template <class EQ, class LT>
void cmp_rows(acu_cmp_op op, int64_t len, bool ls, bool rs, uint8_t *out, EQ &&eq, LT &&lt) {
  auto L = [&](int64_t i) { return ls ? (int64_t)0 : i; };
  auto R = [&](int64_t i) { return rs ? (int64_t)0 : i; };
  switch (op) {
    case ACU_EQ: case ACU_NOT_DISTINCT: collect_bool(len, false, out, [&](int64_t i) { return eq(L(i), R(i)); }); break;
    case ACU_NEQ: case ACU_DISTINCT: collect_bool(len, true, out, [&](int64_t i) { return eq(L(i), R(i)); }); break;
    case ACU_LT: collect_bool(len, false, out, [&](int64_t i) { return lt(false, L(i), R(i)); }); break;
    case ACU_LT_EQ: collect_bool(len, true, out, [&](int64_t i) { return lt(true, R(i), L(i)); }); break;   // !(r < l)
    case ACU_GT: collect_bool(len, false, out, [&](int64_t i) { return lt(true, R(i), L(i)); }); break;      // r < l
    case ACU_GT_EQ: collect_bool(len, true, out, [&](int64_t i) { return lt(false, L(i), R(i)); }); break;  // !(l < r)
  }
}

// `&[u8]` ordering of Rust: lexicographic on unsigned bytes, then length (cmp.rs:783-801 is_eq / is_lt)
inline bool bytes_eq(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb) { return la == lb && (la == 0 || memcmp(a, b, (size_t)la) == 0); }
inline bool bytes_lt(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb) {
  const int64_t n = la < lb ? la : lb;
  const int c = n ? memcmp(a, b, (size_t)n) : 0;
  return c != 0 ? c < 0 : la < lb;
}

// GenericByteViewArray item comparison (cmp.rs:803-898). A view is 16 bytes: length u32, then 12 inline bytes, or
// prefix (4 B) + buffer index u32 + offset u32 (arrow-data/src/byte_view.rs).
struct ViewSide {
  const uint8_t *views;             // 16 bytes per row
  const uint8_t *const *buffers;    // data buffers
  int32_t n_buffers;
};
inline unsigned __int128 view_at(const ViewSide &s, int64_t i) {
  unsigned __int128 v;
  memcpy(&v, s.views + 16 * i, 16);
  return v;
}
inline const uint8_t *view_bytes(const ViewSide &s, int64_t i, uint32_t *len) {  // GenericByteViewArray::value_unchecked
  const uint8_t *v = s.views + 16 * i;
  memcpy(len, v, 4);
  if (*len <= 12) return v + 4;
  uint32_t buf, off;
  memcpy(&buf, v + 8, 4);
  memcpy(&off, v + 12, 4);
  return s.buffers[buf] + off;
}
inline unsigned __int128 bswap128(unsigned __int128 x) {
  const uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
  return ((unsigned __int128)__builtin_bswap64(lo) << 64) | (unsigned __int128)__builtin_bswap64(hi);
}
inline unsigned __int128 inline_key_fast(unsigned __int128 raw) {  // byte_view_array.rs:872-874
  return (bswap128(raw) << 32) | (unsigned __int128)(uint32_t)raw;
}
inline bool view_is_eq(const ViewSide &l, int64_t li, const ViewSide &r, int64_t ri) {  // cmp.rs:810-862
  const unsigned __int128 lv = view_at(l, li), rv = view_at(r, ri);
  if (l.n_buffers == 0 && r.n_buffers == 0) return lv == rv;
  if (lv == rv && (uint32_t)lv <= 12) return true;
  const uint32_t ll = (uint32_t)lv, rl = (uint32_t)rv;
  if (ll != rl) return false;
  if (ll == 0) return true;
  if ((uint32_t)(lv >> 32) != (uint32_t)(rv >> 32)) return false;
  if (ll <= 12) return false;
  uint32_t a, b;
  const uint8_t *pa = view_bytes(l, li, &a), *pb = view_bytes(r, ri, &b);
  return memcmp(pa, pb, ll) == 0;
}
inline bool view_is_lt(const ViewSide &l, int64_t li, const ViewSide &r, int64_t ri) {  // cmp.rs:864-893
  const unsigned __int128 lv = view_at(l, li), rv = view_at(r, ri);
  if (l.n_buffers == 0 && r.n_buffers == 0) return inline_key_fast(lv) < inline_key_fast(rv);
  if ((uint32_t)lv <= 12 && (uint32_t)rv <= 12) return inline_key_fast(lv) < inline_key_fast(rv);
  const uint32_t lp = (uint32_t)(lv >> 32), rp = (uint32_t)(rv >> 32);
  if (lp != rp) return __builtin_bswap32(lp) < __builtin_bswap32(rp);
  uint32_t a, b;
  const uint8_t *pa = view_bytes(l, li, &a), *pb = view_bytes(r, ri, &b);
  return bytes_lt(pa, a, pb, b);
}

// ---- cast: arrow-cast/src/cast/mod.rs:2550-2614 + num-traits 0.2.19 cast -----------------

template <class I, class O>
inline bool num_cast(I v, O *out) {
  if constexpr (std::is_floating_point<O>::value) {
    *out = (O)v;  // int->float and float->float: `as` (RNE / inf on overflow), always Some
    return true;
  } else if constexpr (std::is_floating_point<I>::value) {
    // num-traits float_to_int: accept the exclusive range (MIN-1, MAX+1), truncate toward zero
    if (v != v) return false;
    if constexpr (std::is_signed<O>::value) {
      if constexpr (sizeof(I) > sizeof(O)) {
        const I min_m1 = (I)std::numeric_limits<O>::min() - (I)1, max_p1 = (I)std::numeric_limits<O>::max() + (I)1;
        if (!(v > min_m1 && v < max_p1)) return false;
      } else {
        const I mn = (I)std::numeric_limits<O>::min(), max_p1 = (I)std::numeric_limits<O>::max();
        if (!(v >= mn && v < max_p1)) return false;
      }
    } else {
      if constexpr (sizeof(I) > sizeof(O)) {
        const I max_p1 = (I)std::numeric_limits<O>::max() + (I)1;
        if (!(v > (I)-1 && v < max_p1)) return false;
      } else {
        const I max_p1 = (I)std::numeric_limits<O>::max();
        if (!(v > (I)-1 && v < max_p1)) return false;
      }
    }
    *out = (O)v;
    return true;
  } else {
    // int -> int: Some iff representable
    if constexpr (std::is_signed<I>::value == std::is_signed<O>::value) {
      if (v < std::numeric_limits<O>::min() || v > std::numeric_limits<O>::max()) return false;
    } else if constexpr (std::is_signed<I>::value) {  // signed -> unsigned
      if (v < 0) return false;
      if ((typename std::make_unsigned<I>::type)v > std::numeric_limits<O>::max()) return false;
    } else {  // unsigned -> signed
      if (v > (typename std::make_unsigned<O>::type)std::numeric_limits<O>::max()) return false;
    }
    *out = (O)v;
    return true;
  }
}

const char *dtype_name(acu_dtype t) {  // DataType Display
  static const char *n[] = {"Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16", "UInt32", "UInt64", "Float32", "Float64"};
  return n[t];
}

template <class I, class O>
acu_status cast_typed(int32_t safe, acu_dtype to, const acu_array *a, acu_array_out *out) {
  const I *av = static_cast<const I *>(a->values);
  O *ov = static_cast<O *>(out->values);
  int64_t len = a->len;
  out->len = len;
  memset(ov, 0, (size_t)len * sizeof(O));
  if (safe) {  // numeric_cast = unary_opt (primitive_array.rs:1065-1103): always Some(nulls)
    if (a->validity) clone_nulls(a, len, out);
    else { memset(out->validity, 0xff, acu_bitmap_bytes(len)); out->has_validity = 1; out->null_count = 0; }
    int64_t nc = out->null_count;
    for (int64_t i = 0; i < len; ++i) {
      if (a->validity && !get_bit(a->validity, a->validity_offset + i)) continue;
      O o;
      if (num_cast<I, O>(av[i], &o)) ov[i] = o;
      else { ++nc; out->validity[i >> 3] &= (uint8_t)~(1u << (i & 7)); }
    }
    out->null_count = nc;
    return ACU_OK;
  }
  clone_nulls(a, len, out);  // try_numeric_cast = try_unary
  if (a->validity && out->null_count == len) return ACU_OK;
  for (int64_t i = 0; i < len; ++i) {
    if (a->validity && !get_bit(a->validity, a->validity_offset + i)) continue;
    O o;
    if (!num_cast<I, O>(av[i], &o)) {
      char s[40];
      fmt_native(s, sizeof s, av[i]);
      return fail(ACU_ERR_CAST, i, to_bits(av[i]), 0, 0, "Can't cast value %s to type %s", s, dtype_name(to));
    }
    ov[i] = o;
  }
  return ACU_OK;
}

// ---- aggregate: arrow-arith/src/aggregate.rs ---------------------------------------------

template <class T> struct Limits {
  static T min_total() { if constexpr (std::is_floating_point<T>::value) return from_bits<T>(~0ull); else return std::numeric_limits<T>::min(); }
  static T max_total() {
    if constexpr (std::is_floating_point<T>::value) return from_bits<T>(sizeof(T) == 8 ? 0x7fffffffffffffffull : 0x7fffffffull);
    else return std::numeric_limits<T>::max();
  }
};

template <class T>
struct Acc {  // Sum/Min/MaxAccumulator (aggregate.rs:52-176)
  T v;
  static Acc init(acu_agg_op op) {
    return Acc{op == ACU_SUM ? (T)0 : op == ACU_MIN ? Limits<T>::max_total() : Limits<T>::min_total()};
  }
  inline void accumulate(acu_agg_op op, T x) {
    if (op == ACU_SUM) {
      if constexpr (std::is_floating_point<T>::value) v = v + x;
      else v = (T)((typename std::make_unsigned<T>::type)v + (typename std::make_unsigned<T>::type)x);
    } else if (op == ACU_MIN) {
      v = is_lt(x, v) ? x : v;
    } else {
      v = is_gt(x, v) ? x : v;
    }
  }
  inline void accumulate_nullable(acu_agg_op op, T x, bool valid) { if (valid) accumulate(op, x); }
};

template <class T>
T reduce_lanes(acu_agg_op op, std::vector<Acc<T>> &acc) {  // reduce_accumulators (:178-196)
  size_t len = acc.size();
  while (len >= 2) {
    size_t mid = len / 2;
    for (size_t i = 0; i < mid; ++i) acc[i].accumulate(op, acc[mid + i].v);
    len /= 2;
  }
  return acc[0].v;
}

// aggregate (aggregate.rs:317-366). `vector_bytes` = PREFERRED_VECTOR_SIZE of the
// reference build (:303-311): 16 for baseline x86_64, 32 with AVX, 64 with AVX-512.
template <class T>
acu_status aggregate_typed(acu_agg_op op, const acu_array *a, int32_t vector_bytes, uint64_t *out_bits, int64_t *out_valid) {
  const T *v = static_cast<const T *>(a->values);
  int64_t len = a->len;
  int64_t nc = resolve_null_count(a);
  *out_valid = len - nc;
  *out_bits = 0;
  if (nc == len) return ACU_OK;  // None (also len == 0)
  if (a->validity && nc > 0) {  // aggregate_nullable_lanes (:254-300)
    size_t lanes = (size_t)vector_bytes / sizeof(T);
    if (lanes < 1) lanes = 1;
    if (lanes > 64) lanes = 64;
    std::vector<Acc<T>> acc(lanes, Acc<T>::init(op));
    int64_t full = len / 64 * 64;
    for (int64_t c = 0; c < full; c += 64) {
      uint64_t val = load_bits(a->validity, a->validity_offset + c, a->validity_offset + len);
      for (int64_t k = 0; k < 64; k += (int64_t)lanes) {
        for (size_t i = 0; i < lanes; ++i) acc[i].accumulate_nullable(op, v[c + k + (int64_t)i], (val >> i) & 1);
        val = lanes < 64 ? val >> lanes : 0;
      }
    }
    int64_t rem = len - full;
    if (rem) {
      uint64_t val = load_bits(a->validity, a->validity_offset + full, a->validity_offset + len);
      int64_t k = 0;
      for (; k + (int64_t)lanes <= rem; k += (int64_t)lanes) {
        for (size_t i = 0; i < lanes; ++i) acc[i].accumulate_nullable(op, v[full + k + (int64_t)i], (val >> i) & 1);
        val = lanes < 64 ? val >> lanes : 0;
      }
      for (int64_t i = 0; k + i < rem; ++i) acc[(size_t)i].accumulate_nullable(op, v[full + k + i], (val >> i) & 1);
    }
    *out_bits = to_bits(reduce_lanes(op, acc));
    return ACU_OK;
  }
  size_t lanes = std::is_floating_point<T>::value ? (size_t)vector_bytes * 2 / sizeof(T) : 1;
  if (lanes > 64) lanes = 64;
  if (lanes <= 1) {  // aggregate_nonnull_simple (:222-231)
    Acc<T> acc = Acc<T>::init(op);
    for (int64_t i = 0; i < len; ++i) acc.accumulate(op, v[i]);
    *out_bits = to_bits(acc.v);
    return ACU_OK;
  }
  std::vector<Acc<T>> acc(lanes, Acc<T>::init(op));  // aggregate_nonnull_lanes (:234-251)
  int64_t i = 0;
  for (; i + (int64_t)lanes <= len; i += (int64_t)lanes)
    for (size_t k = 0; k < lanes; ++k) acc[k].accumulate(op, v[i + (int64_t)k]);
  for (int64_t k = 0; i + k < len; ++k) acc[(size_t)k].accumulate(op, v[i + k]);
  *out_bits = to_bits(reduce_lanes(op, acc));
  return ACU_OK;
}

// ---- synthetic inputs; identical generator in csrc/generate.cu ---------
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

#define DISPATCH_DTYPE(dt, F, ...)                            \
  switch (dt) {                                               \
    case ACU_I8: return F<int8_t>(__VA_ARGS__);               \
    case ACU_I16: return F<int16_t>(__VA_ARGS__);             \
    case ACU_I32: return F<int32_t>(__VA_ARGS__);             \
    case ACU_I64: return F<int64_t>(__VA_ARGS__);             \
    case ACU_U8: return F<uint8_t>(__VA_ARGS__);              \
    case ACU_U16: return F<uint16_t>(__VA_ARGS__);            \
    case ACU_U32: return F<uint32_t>(__VA_ARGS__);            \
    case ACU_U64: return F<uint64_t>(__VA_ARGS__);            \
    case ACU_F32: return F<float>(__VA_ARGS__);               \
    case ACU_F64: return F<double>(__VA_ARGS__);              \
  }

template <class I>
acu_status cast_from(acu_dtype to, int32_t safe, const acu_array *a, acu_array_out *out) {
  switch (to) {
    case ACU_I8: return cast_typed<I, int8_t>(safe, to, a, out);
    case ACU_I16: return cast_typed<I, int16_t>(safe, to, a, out);
    case ACU_I32: return cast_typed<I, int32_t>(safe, to, a, out);
    case ACU_I64: return cast_typed<I, int64_t>(safe, to, a, out);
    case ACU_U8: return cast_typed<I, uint8_t>(safe, to, a, out);
    case ACU_U16: return cast_typed<I, uint16_t>(safe, to, a, out);
    case ACU_U32: return cast_typed<I, uint32_t>(safe, to, a, out);
    case ACU_U64: return cast_typed<I, uint64_t>(safe, to, a, out);
    case ACU_F32: return cast_typed<I, float>(safe, to, a, out);
    case ACU_F64: return cast_typed<I, double>(safe, to, a, out);
  }
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

}  // namespace

// sum_checked (arrow-arith/src/aggregate.rs:897-937): try_fold(add_checked) over the valid values in order;
// the error is add_checked's (arrow-array/src/arithmetic.rs:163-170). Floats: add_checked = plain add, in order.
template <class T>
acu_status sum_checked_typed(const acu_array *a, uint64_t *out_bits, int64_t *out_valid) {
  const T *v = static_cast<const T *>(a->values);
  const int64_t len = a->len, nc = resolve_null_count(a);
  *out_valid = len - nc;
  *out_bits = 0;
  if (nc == len) return ACU_OK;  // Ok(None)
  T acc = T();
  for (int64_t i = 0; i < len; ++i) {
    if (a->validity && !get_bit(a->validity, a->validity_offset + i)) continue;
    if constexpr (std::is_floating_point<T>::value) {
      acc = acc + v[i];
    } else {
      T r;
      if (__builtin_add_overflow(acc, v[i], &r)) {
        char ls[40], rs[40];
        fmt_native(ls, sizeof ls, acc);
        fmt_native(rs, sizeof rs, v[i]);
        return fail(ACU_ERR_ARITHMETIC_OVERFLOW, i, to_bits(acc), to_bits(v[i]), 0, "Overflow happened on: %s + %s", ls, rs);
      }
      acc = r;
    }
  }
  *out_bits = to_bits(acc);
  return ACU_OK;
}

extern "C" {

const acu_error_detail *orc_last_error(void) { return &g_err; }

acu_status orc_bitmap_count(const uint8_t *bits, int64_t offset, const uint8_t *validity,
                            int64_t validity_offset, int64_t len, int64_t *out_count) {
  int64_t c = 0;
  for (int64_t i = 0; i < len; i += 64) {
    uint64_t v = load_bits(bits, offset + i, offset + len);
    if (validity) v &= load_bits(validity, validity_offset + i, validity_offset + len);
    c += __builtin_popcountll(v);
  }
  *out_count = c;
  return ACU_OK;
}

// filter() for primitive arrays (filter.rs:201-213 -> filter_array :535-546 ->
// filter_primitive :773-788). out_count / out_strategy expose FilterPredicate internals.
acu_status orc_filter_primitive(const acu_array *predicate, int32_t elem_bytes,
                                const acu_array *values, acu_array_out *out, int64_t *out_count,
                                int32_t *out_strategy) {
  Predicate p;
  build_predicate(predicate, &p);
  if (out_count) *out_count = p.count;
  if (out_strategy) *out_strategy = p.strategy;
  if (acu_status st = check_filter_len(p, values->len)) return st;
  out->len = p.count;
  out->has_validity = 0;
  out->null_count = 0;
  const uint8_t *src = static_cast<const uint8_t *>(values->values);
  uint8_t *dst = static_cast<uint8_t *>(out->values);
  const size_t w = (size_t)elem_bytes;
  switch (p.strategy) {
    case ACU_FILTER_NONE: return ACU_OK;
    case ACU_FILTER_ALL:
      memcpy(dst, src, (size_t)p.count * w);
      slice_nulls(values, p.count, out);
      return ACU_OK;
    case ACU_FILTER_SLICES: {  // filter_native SlicesIterator arm (:740-747)
      size_t k = 0;
      for_each_slice(p, [&](int64_t s, int64_t e) {
        memcpy(dst + k * w, src + (size_t)s * w, (size_t)(e - s) * w);
        k += (size_t)(e - s);
      });
      break;
    }
    default: {  // IndexIterator arm (:756-763)
      size_t k = 0;
      if (w == 8) {
        const uint64_t *s8 = reinterpret_cast<const uint64_t *>(src);
        uint64_t *d8 = reinterpret_cast<uint64_t *>(dst);
        for_each_index(p, [&](int64_t i) { d8[k++] = s8[i]; });
      } else if (w == 4) {
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
        for_each_index(p, [&](int64_t i) { d4[k++] = s4[i]; });
      } else {
        for_each_index(p, [&](int64_t i) { memcpy(dst + (k++) * w, src + (size_t)i * w, w); });
      }
    }
  }
  filter_nulls(p, values, out);
  return ACU_OK;
}

// filter_boolean (filter.rs:723-729)
acu_status orc_filter_boolean(const acu_array *predicate, const acu_array *values,
                              acu_array_out *out, int64_t *out_count) {
  Predicate p;
  build_predicate(predicate, &p);
  if (out_count) *out_count = p.count;
  if (acu_status st = check_filter_len(p, values->len)) return st;
  out->len = p.count;
  out->has_validity = 0;
  out->null_count = 0;
  const uint8_t *src = static_cast<const uint8_t *>(values->values);
  uint8_t *dst = static_cast<uint8_t *>(out->values);
  switch (p.strategy) {
    case ACU_FILTER_NONE: return ACU_OK;
    case ACU_FILTER_ALL:
      memset(dst, 0, acu_bitmap_bytes(p.count));
      for (int64_t i = 0; i < p.count; ++i)
        if (get_bit(src, values->values_offset + i)) set_bit(dst, i);
      slice_nulls(values, p.count, out);
      return ACU_OK;
    default: filter_bits(p, src, values->values_offset, dst);
  }
  filter_nulls(p, values, out);
  return ACU_OK;
}

// filter_bytes (filter.rs:893-928): offsets pass then bytes pass; null slots with
// non-zero length are copied too (:891-892).
acu_status orc_filter_bytes(const acu_array *predicate, int32_t offset_bytes, const void *offsets,
                            const uint8_t *data, const acu_array *nulls_of, void *out_offsets,
                            uint8_t *out_data, int64_t out_data_capacity, int64_t *out_data_len,
                            acu_array_out *out_nulls, int64_t *out_count) {
  Predicate p;
  build_predicate(predicate, &p);
  if (out_count) *out_count = p.count;
  if (acu_status st = check_filter_len(p, nulls_of->len)) return st;
  out_nulls->len = p.count;
  out_nulls->has_validity = 0;
  out_nulls->null_count = 0;
  auto off = [&](int64_t i) -> int64_t {
    return offset_bytes == 4 ? (int64_t) static_cast<const int32_t *>(offsets)[i]
                             : static_cast<const int64_t *>(offsets)[i];
  };
  auto put = [&](int64_t k, int64_t v) {
    if (offset_bytes == 4) static_cast<int32_t *>(out_offsets)[k] = (int32_t)v;
    else static_cast<int64_t *>(out_offsets)[k] = v;
  };
  put(0, 0);
  *out_data_len = 0;
  if (p.strategy == ACU_FILTER_NONE) return ACU_OK;
  int64_t cur = 0, k = 0;
  auto row = [&](int64_t i) {
    int64_t s = off(i), e = off(i + 1);
    if (out_data && cur + (e - s) <= out_data_capacity) memcpy(out_data + cur, data + s, (size_t)(e - s));
    cur += e - s;
    put(++k, cur);
  };
  if (p.strategy == ACU_FILTER_ALL) {
    for (int64_t i = 0; i < p.count; ++i) row(i);
    *out_data_len = cur;
    slice_nulls(nulls_of, p.count, out_nulls);
    return ACU_OK;
  }
  if (p.strategy == ACU_FILTER_SLICES)
    for_each_slice(p, [&](int64_t s, int64_t e) { for (int64_t i = s; i < e; ++i) row(i); });
  else
    for_each_index(p, row);
  *out_data_len = cur;
  filter_nulls(p, nulls_of, out_nulls);
  return ACU_OK;
}

// take() for primitive arrays (take.rs:89-105 -> take_impl :212-219 -> take_primitive :405-416)
acu_status orc_take_primitive(int32_t elem_bytes, const acu_array *values, const acu_array *indices,
                              acu_dtype index_dtype, int32_t check_bounds_opt, acu_array_out *out) {
  if (!is_int_dtype(index_dtype))
    return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Take only supported for integers, got %s", dtype_name(index_dtype));
  if (check_bounds_opt)
    if (acu_status st = check_bounds(values->len, indices, index_dtype)) return st;
  int64_t m = indices->len;
  out->len = m;
  out->has_validity = 0;
  out->null_count = 0;
  if (m == 0) return ACU_OK;
  if (acu_status st = check_panic(values->len, indices, index_dtype)) return st;
  const uint8_t *src = static_cast<const uint8_t *>(values->values);
  uint8_t *dst = static_cast<uint8_t *>(out->values);
  const size_t w = (size_t)elem_bytes;
  const uint64_t n = (uint64_t)values->len;
  // take_native (take.rs:433-457): in-bounds => the stored value (even at null index
  // slots); out-of-bounds (only possible at a null index after check_panic) => default
  if (w == 8 && (index_dtype == ACU_U32 || index_dtype == ACU_I32)) {
    const uint64_t *s8 = reinterpret_cast<const uint64_t *>(src);
    uint64_t *d8 = reinterpret_cast<uint64_t *>(dst);
    const uint32_t *ix = static_cast<const uint32_t *>(indices->values);
    for (int64_t j = 0; j < m; ++j) d8[j] = ix[j] < n ? s8[ix[j]] : 0;
  } else {
    for (int64_t j = 0; j < m; ++j) {
      uint64_t ix = load_index(indices->values, index_dtype, j);
      if (ix < n) memcpy(dst + (size_t)j * w, src + (size_t)ix * w, w);
      else memset(dst + (size_t)j * w, 0, w);
    }
  }
  take_nulls(values, indices, index_dtype, out);
  return ACU_OK;
}

// take_boolean (take.rs:489-496)
acu_status orc_take_boolean(const acu_array *values, const acu_array *indices, acu_dtype index_dtype,
                            int32_t check_bounds_opt, acu_array_out *out) {
  if (check_bounds_opt)
    if (acu_status st = check_bounds(values->len, indices, index_dtype)) return st;
  int64_t m = indices->len;
  out->len = m;
  out->has_validity = 0;
  out->null_count = 0;
  if (m == 0) return ACU_OK;
  if (acu_status st = check_panic(values->len, indices, index_dtype)) return st;
  bool ind_nulls = indices->validity && resolve_null_count(indices) > 0;
  uint8_t *dst = static_cast<uint8_t *>(out->values);
  memset(dst, 0, acu_bitmap_bytes(m));
  const uint8_t *src = static_cast<const uint8_t *>(values->values);
  for (int64_t j = 0; j < m; ++j) {  // take_bits (take.rs:460-486)
    if (ind_nulls && !get_bit(indices->validity, indices->validity_offset + j)) continue;
    uint64_t ix = load_index(indices->values, index_dtype, j);
    if (get_bit(src, values->values_offset + (int64_t)ix)) set_bit(dst, j);
  }
  take_nulls(values, indices, index_dtype, out);
  return ACU_OK;
}

// take_bytes (take.rs:499-627)
acu_status orc_take_bytes(int32_t offset_bytes, const void *offsets, const uint8_t *data,
                          const acu_array *nulls_of, const acu_array *indices, acu_dtype index_dtype,
                          int32_t check_bounds_opt, void *out_offsets, uint8_t *out_data,
                          int64_t out_data_capacity, int64_t *out_data_len, acu_array_out *out_nulls) {
  if (check_bounds_opt)
    if (acu_status st = check_bounds(nulls_of->len, indices, index_dtype)) return st;
  int64_t m = indices->len;
  out_nulls->len = m;
  out_nulls->has_validity = 0;
  out_nulls->null_count = 0;
  *out_data_len = 0;
  auto off = [&](int64_t i) -> int64_t {
    return offset_bytes == 4 ? (int64_t) static_cast<const int32_t *>(offsets)[i]
                             : static_cast<const int64_t *>(offsets)[i];
  };
  auto put = [&](int64_t k, int64_t v) {
    if (offset_bytes == 4) static_cast<int32_t *>(out_offsets)[k] = (int32_t)v;
    else static_cast<int64_t *>(out_offsets)[k] = v;
  };
  put(0, 0);
  if (m == 0) return ACU_OK;  // take_impl: new_empty_array (offsets = [0])
  // take_nulls runs first (take.rs:509): with nulls in the values it is take_bits over the validity, whose
  // bounds-checked BooleanBuffer::value panics on an out-of-bounds index at a valid slot (take.rs:472, :482)
  if (nulls_of->validity && resolve_null_count(nulls_of) > 0)
    if (acu_status st = check_panic(nulls_of->len, indices, index_dtype)) return st;
  take_nulls(nulls_of, indices, index_dtype, out_nulls);
  const int64_t limit = offset_bytes == 4 ? (int64_t)INT32_MAX : INT64_MAX;
  int64_t cap = 0;
  const bool out_has_nulls = out_nulls->has_validity && out_nulls->null_count > 0;
  const uint64_t n = (uint64_t)nulls_of->len;
  for (int64_t j = 0; j < m; ++j) {
    bool valid = !out_has_nulls || get_bit(out_nulls->validity, j);
    if (valid) {
      uint64_t ix = load_index(indices->values, index_dtype, j);
      if (ix >= n)  // bounds-checked slice index `input_offsets[index]` panics
        return fail(ACU_ERR_PANIC_OUT_OF_BOUNDS, j, ix, 0, n, "Out-of-bounds index %" PRIu64, ix);
      int64_t s = off((int64_t)ix), e = off((int64_t)ix + 1);
      if (out_data && cap + (e - s) <= out_data_capacity) memcpy(out_data + cap, data + s, (size_t)(e - s));
      cap += e - s;
      if (cap > limit)  // T::Offset::from_usize(capacity) (take.rs:520-523, :561-574)
        return fail(ACU_ERR_OFFSET_OVERFLOW, j, 0, 0, (uint64_t)cap, "%" PRId64, cap);
    }
    put(j + 1, cap);
  }
  *out_data_len = cap;
  return ACU_OK;
}

acu_status orc_arith(acu_dtype dtype, acu_arith_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  DISPATCH_DTYPE(dtype, arith_typed, op, a, b, out)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

acu_status orc_neg(acu_dtype dtype, int32_t checked, const acu_array *a, acu_array_out *out) {
  if (checked && (dtype == ACU_U8 || dtype == ACU_U16 || dtype == ACU_U32 || dtype == ACU_U64))
    return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "Invalid arithmetic operation: !%s", dtype_name(dtype));
  DISPATCH_DTYPE(dtype, neg_typed, checked, a, out)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

// arrow-arith/src/boolean.rs: and :256, or :273, and_not :291 (binary_boolean_kernel :224-243), and_kleene :60-124,
// or_kleene :156-222, not :310-314, is_null :327-334, is_not_null :347-354 — restated row by row with Option<bool>
// semantics rather than the reference's word-level bit formulas, so that the formulas themselves are checked.
acu_status orc_boolean(acu_bool_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  const bool binary = op <= ACU_BOOL_OR_KLEENE;
  if (binary && a->len != b->len)
    return fail(ACU_ERR_COMPUTE, -1, 0, 0, 0, "Cannot perform bitwise operation on arrays of different length");
  const int64_t n = a->len;
  out->len = n;
  out->null_count = 0;
  const bool nulls_out = op == ACU_BOOL_NOT ? a->validity != nullptr : binary ? (a->validity != nullptr || b->validity != nullptr) : false;
  out->has_validity = nulls_out;
  uint8_t *ov = static_cast<uint8_t *>(out->values);
  memset(ov, 0, acu_bitmap_bytes(n));
  if (nulls_out) memset(out->validity, 0, acu_bitmap_bytes(n));
  int64_t valid_rows = 0;
  for (int64_t i = 0; i < n; ++i) {
    const bool a_valid = !a->validity || get_bit(a->validity, a->validity_offset + i);
    const bool av = (op == ACU_BOOL_IS_NULL || op == ACU_BOOL_IS_NOT_NULL) ? false : get_bit(static_cast<const uint8_t *>(a->values), a->values_offset + i);
    const bool b_valid = !binary || !b->validity || get_bit(b->validity, b->validity_offset + i);
    const bool bv = binary ? get_bit(static_cast<const uint8_t *>(b->values), b->values_offset + i) : false;
    bool v = false, valid = true;
    switch (op) {
      case ACU_BOOL_AND: v = av && bv; valid = a_valid && b_valid; break;          // the raw bits combine even under nulls
      case ACU_BOOL_OR: v = av || bv; valid = a_valid && b_valid; break;
      case ACU_BOOL_AND_NOT: v = av && !bv; valid = a_valid && b_valid; break;
      case ACU_BOOL_AND_KLEENE:
        v = av && bv;
        // known false on either side decides; otherwise both must be known
        valid = (a_valid && b_valid) || (a_valid && !av) || (b_valid && !bv);
        break;
      case ACU_BOOL_OR_KLEENE:
        v = av || bv;
        valid = (a_valid && b_valid) || (a_valid && av) || (b_valid && bv);
        break;
      case ACU_BOOL_NOT: v = !av; valid = a_valid; break;
      case ACU_BOOL_IS_NULL: v = !a_valid; break;
      default: v = a_valid; break;
    }
    if (v) set_bit(ov, i);
    if (nulls_out && valid) { set_bit(out->validity, i); ++valid_rows; }
  }
  if (nulls_out) out->null_count = n - valid_rows;
  return ACU_OK;
}

acu_status orc_cmp(acu_dtype dtype, acu_cmp_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  DISPATCH_DTYPE(dtype, cmp_typed, op, a, b, out)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

acu_status orc_cast_numeric(acu_dtype from, acu_dtype to, int32_t safe, const acu_array *a, acu_array_out *out) {
  DISPATCH_DTYPE(from, cast_from, to, safe, a, out)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

// vector_bytes: PREFERRED_VECTOR_SIZE of the reference build being modelled (16/32/64).
acu_status orc_aggregate(acu_dtype dtype, acu_agg_op op, const acu_array *a, int32_t vector_bytes,
                         uint64_t *out_bits, int64_t *out_valid_count) {
  DISPATCH_DTYPE(dtype, aggregate_typed, op, a, vector_bytes, out_bits, out_valid_count)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

acu_status orc_sum_checked(acu_dtype dtype, const acu_array *a, uint64_t *out_bits, int64_t *out_valid_count) {
  DISPATCH_DTYPE(dtype, sum_checked_typed, a, out_bits, out_valid_count)
  return ACU_ERR_NOT_YET_IMPLEMENTED;
}

acu_status orc_generate_values(int32_t kind, uint64_t seed, int64_t first_row, uint64_t param, void *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t r = splitmix64(seed ^ (uint64_t)(first_row + i));
    switch (kind) {
      case 0: static_cast<uint64_t *>(out)[i] = r; break;
      case 1: static_cast<int64_t *>(out)[i] = (int64_t)(r >> 2) - ((int64_t)1 << 61); break;
      case 2: static_cast<double *>(out)[i] = (double)(r >> 11) * (2.0e6 / 9007199254740992.0) - 1.0e6; break;
      case 3: static_cast<uint32_t *>(out)[i] = (uint32_t)(((unsigned __int128)r * param) >> 64); break;
      case 4: static_cast<int32_t *>(out)[i] = (int32_t)(((unsigned __int128)r * param) >> 64); break;
      default: return ACU_ERR_INVALID_ARGUMENT;
    }
  }
  return ACU_OK;
}

acu_status orc_generate_bits(uint64_t seed, int64_t first_row, double p, uint8_t *out_bits, int64_t n) {
  uint64_t thr = p >= 1.0 ? ~0ull : (uint64_t)(p * 18446744073709551616.0);
  memset(out_bits, 0, acu_bitmap_bytes(n));
  for (int64_t i = 0; i < n; ++i)
    if (p >= 1.0 || splitmix64(seed ^ (uint64_t)(first_row + i)) < thr) set_bit(out_bits, i);
  return ACU_OK;
}

// nullif (arrow-select/src/nullif.rs:44-113): only the validity changes; `out->validity` receives
// left.nulls & !(right.values & right.nulls) and the builder drops it when it holds no nulls
// (arrow-data/src/data.rs:2238-2251 `.filter(|b| b.null_count() != 0)`).
acu_status orc_nullif(const acu_array *left, const acu_array *right, acu_array_out *out) {
  if (left->len != right->len)  // :47-51
    return fail(ACU_ERR_COMPUTE, -1, 0, 0, 0, "Cannot perform comparison operation on arrays of different length");
  const int64_t len = left->len;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  if (len == 0) return ACU_OK;  // :54-56
  const uint8_t *rv = static_cast<const uint8_t *>(right->values);
  int64_t valid = 0;
  for (int64_t i = 0, w = 0; i < len; i += 64, ++w) {
    uint64_t r = load_bits(rv, right->values_offset + i, right->values_offset + len);
    if (right->validity) r &= load_bits(right->validity, right->validity_offset + i, right->validity_offset + len);  // :70-73
    uint64_t t;
    if (left->validity) {  // bitwise_bin_op_helper(l & !r) (:78-92)
      t = load_bits(left->validity, left->validity_offset + i, left->validity_offset + len) & ~r;
    } else {               // bitwise_unary_op_helper(!r) (:94-103); bits past len are not counted
      t = ~r;
    }
    int64_t n = len - i;
    if (n < 64) t &= (~0ull) >> (64 - n);
    valid += __builtin_popcountll(t);
    memcpy(out->validity + 8 * w, &t, 8);
  }
  out->null_count = len - valid;
  out->has_validity = out->null_count > 0;
  return ACU_OK;
}

// zip (arrow-select/src/zip.rs:99-226) for fixed-width values of `elem_bytes` bytes. Array path: MutableArrayData extends
// runs of truthy (SlicesIterator over the null-cleaned mask) and fills the gaps with falsy (:156-226); values are copied
// blindly, validity is tracked iff some input has nulls (arrow-data/src/transform/mod.rs:470) and dropped when the result has
// none (:927-936). Both sides scalar: PrimitiveScalarImpl::create_output (:392-440).
acu_status orc_zip(int32_t elem_bytes, const acu_array *mask, const acu_array *truthy, const acu_array *falsy, acu_array_out *out) {
  const bool ts = truthy->is_scalar != 0, fs = falsy->is_scalar != 0;
  if (ts && truthy->len != 1) return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "scalar arrays must have 1 element");
  if (!ts && truthy->len != mask->len) return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "all arrays should have the same length");
  if (fs && falsy->len != 1) return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "scalar arrays must have 1 element");
  if (!fs && falsy->len != mask->len) return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "all arrays should have the same length");
  const int64_t len = mask->len;
  const size_t w = (size_t)elem_bytes;
  out->len = len;
  out->has_validity = 0;
  out->null_count = 0;
  if (len == 0) return ACU_OK;
  // maybe_prep_null_mask_filter (:662-671)
  Predicate p;
  {
    const uint8_t *vals = static_cast<const uint8_t *>(mask->values);
    const bool mask_nulls = mask->validity && resolve_null_count(mask) > 0;
    p.len = len;
    p.mask.assign((size_t)((len + 63) / 64) + 1, 0);
    for (int64_t i = 0, k = 0; i < len; i += 64, ++k) {
      uint64_t v = load_bits(vals, mask->values_offset + i, mask->values_offset + len);
      if (mask_nulls) v &= load_bits(mask->validity, mask->validity_offset + i, mask->validity_offset + len);
      p.mask[k] = v;
    }
  }
  const uint8_t *tv = static_cast<const uint8_t *>(truthy->values), *fv = static_cast<const uint8_t *>(falsy->values);
  uint8_t *dst = static_cast<uint8_t *>(out->values);
  auto is_null0 = [](const acu_array *a) { return a->validity && !get_bit(a->validity, a->validity_offset); };
  if (ts && fs) {
    const bool tn = is_null0(truthy), fn = is_null0(falsy);
    const uint8_t *pb = p.bytes();
    if (!tn && !fn) {
      for (int64_t i = 0; i < len; ++i) memcpy(dst + (size_t)i * w, get_bit(pb, i) ? tv : fv, w);
      return ACU_OK;  // nulls: None
    }
    memset(out->validity, 0, acu_bitmap_bytes(len));
    if (!tn || !fn) {  // vec![value; len] + NullBuffer::new(predicate | !predicate)
      const uint8_t *val = !tn ? tv : fv;
      for (int64_t i = 0; i < len; ++i) memcpy(dst + (size_t)i * w, val, w);
      for (int64_t i = 0; i < len; ++i)
        if (get_bit(pb, i) == !tn) set_bit(out->validity, i);
    } else {
      memset(dst, 0, (size_t)len * w);  // T::default_value(), NullBuffer::new_null
    }
    out->has_validity = 1;  // PrimitiveArray::try_new(scalars, Some(nulls)) keeps the buffer even without nulls
    out->null_count = len - count_bits(out->validity, 0, len);
    return ACU_OK;
  }
  const bool use_nulls = resolve_null_count(truthy) > 0 || resolve_null_count(falsy) > 0;
  if (use_nulls) memset(out->validity, 0, acu_bitmap_bytes(len));
  auto extend = [&](const acu_array *src, const uint8_t *sv, bool scalar, int64_t start, int64_t end) {  // try_extend(idx, start, end) / one by one for scalars
    for (int64_t i = start; i < end; ++i) {
      const int64_t j = scalar ? 0 : i;
      memcpy(dst + (size_t)i * w, sv + (size_t)j * w, w);
      if (use_nulls && (!src->validity || get_bit(src->validity, src->validity_offset + j))) set_bit(out->validity, i);
    }
  };
  int64_t filled = 0;
  for_each_slice(p, [&](int64_t s, int64_t e) {
    if (s > filled) extend(falsy, fv, fs, filled, s);
    extend(truthy, tv, ts, s, e);
    filled = e;
  });
  if (filled < len) extend(falsy, fv, fs, filled, len);
  if (use_nulls) {
    const int64_t nc = len - count_bits(out->validity, 0, len);
    if (nc > 0) { out->has_validity = 1; out->null_count = nc; }
  }
  return ACU_OK;
}

// cmp::{eq..not_distinct} on GenericByteArray operands (Utf8 / Binary with i32 offsets, Large* with i64): cmp.rs:220-382 with
// ArrayOrd for &GenericByteArray (:783-801). *_nulls carry len / validity / is_scalar of each side.
acu_status orc_cmp_bytes(int32_t offset_bytes, acu_cmp_op op, const void *l_offsets, const uint8_t *l_data, const acu_array *l_nulls,
                         const void *r_offsets, const uint8_t *r_data, const acu_array *r_nulls, acu_array_out *out) {
  auto off = [&](const void *o, int64_t i) -> int64_t {
    return offset_bytes == 4 ? (int64_t)static_cast<const int32_t *>(o)[i] : static_cast<const int64_t *>(o)[i];
  };
  const bool ls = l_nulls->is_scalar != 0, rs = r_nulls->is_scalar != 0;
  return cmp_generic(op, l_nulls, r_nulls, out, [&](int64_t len, uint8_t *ov) {
    cmp_rows(op, len, ls, rs, ov,
             [&](int64_t i, int64_t j) { return bytes_eq(l_data + off(l_offsets, i), off(l_offsets, i + 1) - off(l_offsets, i),
                                                         r_data + off(r_offsets, j), off(r_offsets, j + 1) - off(r_offsets, j)); },
             [&](bool swapped, int64_t i, int64_t j) {  // swapped: i indexes the RIGHT operand
               const void *ao = swapped ? r_offsets : l_offsets, *bo = swapped ? l_offsets : r_offsets;
               const uint8_t *ad = swapped ? r_data : l_data, *bd = swapped ? l_data : r_data;
               return bytes_lt(ad + off(ao, i), off(ao, i + 1) - off(ao, i), bd + off(bo, j), off(bo, j + 1) - off(bo, j));
             });
  });
}

// cmp on GenericByteViewArray operands (Utf8View / BinaryView): cmp.rs:803-898, plus the inline-constant fast path
// eq_inline_scalar (:405-435), which yields the same bits as the generic path for valid views.
acu_status orc_cmp_byte_view(acu_cmp_op op, const void *l_views, const uint8_t *const *l_buffers, int32_t l_n_buffers,
                             const acu_array *l_nulls, const void *r_views, const uint8_t *const *r_buffers, int32_t r_n_buffers,
                             const acu_array *r_nulls, acu_array_out *out) {
  const ViewSide L{static_cast<const uint8_t *>(l_views), l_buffers, l_n_buffers}, R{static_cast<const uint8_t *>(r_views), r_buffers, r_n_buffers};
  const bool ls = l_nulls->is_scalar != 0, rs = r_nulls->is_scalar != 0;
  // eq_inline_scalar (cmp.rs:282-300, :405-435): == / != of an array against a non-null constant of <= 4 bytes compares the
  // low 64 bits of every view (length + prefix) under a mask; nulls = the array side's (when it has any)
  if ((op == ACU_EQ || op == ACU_NEQ) && ls != rs) {
    const acu_array *arr = ls ? r_nulls : l_nulls, *sc = ls ? l_nulls : r_nulls;
    const ViewSide &A = ls ? R : L, &S = ls ? L : R;
    if (sc->len >= 1 && !(sc->validity && resolve_null_count(sc) > 0)) {
      const unsigned __int128 nv = view_at(S, 0);
      const uint32_t needle_len = (uint32_t)nv;
      if (needle_len <= 4) {
        const uint64_t significant = ~0ull >> (32 - needle_len * 8);
        const uint64_t needle = (uint64_t)nv & significant;
        const int64_t len = arr->len;
        out->len = len;
        out->has_validity = 0;
        out->null_count = 0;
        collect_bool(len, op == ACU_NEQ, static_cast<uint8_t *>(out->values), [&](int64_t i) { return ((uint64_t)view_at(A, i) & significant) == needle; });
        if (arr->validity && resolve_null_count(arr) > 0) clone_nulls(arr, len, out);
        return ACU_OK;
      }
    }
  }
  return cmp_generic(op, l_nulls, r_nulls, out, [&](int64_t len, uint8_t *ov) {
    cmp_rows(op, len, ls, rs, ov, [&](int64_t i, int64_t j) { return view_is_eq(L, i, R, j); },
             [&](bool swapped, int64_t i, int64_t j) { return swapped ? view_is_lt(R, i, L, j) : view_is_lt(L, i, R, j); });
  });
}

// concat (arrow-select/src/concat.rs:495-577) over host columns: concat_primitives / concat_boolean / concat_bytes are
// builder.append_array per input (:334-368). acu_column as in include/arrow_cuda.h with HOST pointers.
acu_status orc_concat(int32_t n, const acu_column *cols, acu_column_out *out) {
  if (n <= 0) return fail(ACU_ERR_COMPUTE, -1, 0, 0, 0, "concat requires input of at least one array");  // :496-499
  for (int i = 1; i < n; ++i)
    if (cols[i].kind != cols[0].kind || cols[i].width != cols[0].width)
      return fail(ACU_ERR_INVALID_ARGUMENT, i, 0, 0, 0,
                  "It is not possible to concatenate arrays of different data types (kind %d width %d, kind %d width %d).",
                  cols[0].kind, cols[0].width, cols[i].kind, cols[i].width);
  const int kind = cols[0].kind, w = cols[0].width;
  int64_t total = 0;
  bool materialized = false;  // NullBufferBuilder::append_buffer materialises on the first buffer with nulls (null.rs:209-218)
  for (int i = 0; i < n; ++i) {
    total += cols[i].array.len;
    if (cols[i].array.validity && resolve_null_count(&cols[i].array) > 0) materialized = true;
  }
  const bool single = n == 1;  // array.slice(0, len): buffers shared, NullBuffer kept as it is (:500-503)
  const bool keep = single && cols[0].array.validity != nullptr;
  out->array.len = total;
  out->array.has_validity = 0;
  out->array.null_count = 0;
  out->data_len = 0;
  if (materialized || keep) memset(out->array.validity, 0, acu_bitmap_bytes(total));
  if (kind == ACU_COL_BOOLEAN) memset(out->array.values, 0, acu_bitmap_bytes(total));
  int64_t row = 0, bytes = 0;
  if (kind == ACU_COL_BYTES) memset(out->array.values, 0, (size_t)w);
  for (int i = 0; i < n; ++i) {
    const acu_column &c = cols[i];
    const int64_t len = c.array.len;
    if (kind == ACU_COL_PRIMITIVE) {
      memcpy(static_cast<uint8_t *>(out->array.values) + (size_t)row * w, c.array.values, (size_t)len * w);  // extend_from_slice(values)
    } else if (kind == ACU_COL_BOOLEAN) {
      const uint8_t *src = static_cast<const uint8_t *>(c.array.values);
      for (int64_t k = 0; k < len; ++k)
        if (get_bit(src, c.array.values_offset + k)) set_bit(static_cast<uint8_t *>(out->array.values), row + k);
    } else if (len > 0) {  // GenericByteBuilder::append_array (generic_bytes_builder.rs:169-206)
      auto off = [&](int64_t k) -> int64_t { return w == 4 ? (int64_t)static_cast<const int32_t *>(c.array.values)[k] : static_cast<const int64_t *>(c.array.values)[k]; };
      const int64_t first = off(0), last = off(len);
      const int64_t limit = w == 4 ? (int64_t)INT32_MAX : INT64_MAX;
      if (bytes != first) {  // shifting all the offsets: checked_add(shift, last)
        const int64_t shift = bytes - first;
        if (shift + last > limit || (w == 4 && shift + last < (int64_t)INT32_MIN))
          return fail(ACU_ERR_OFFSET_OVERFLOW, -1, 0, 0, (uint64_t)(shift + last), "%" PRId64, shift + last);
      }
      for (int64_t k = 1; k <= len; ++k) {
        const int64_t v = bytes + (off(k) - first);
        if (w == 4) static_cast<int32_t *>(out->array.values)[row + k] = (int32_t)v;
        else static_cast<int64_t *>(out->array.values)[row + k] = v;
      }
      if (bytes + (last - first) > out->data_capacity)
        return fail(ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, (uint64_t)(bytes + (last - first)), "output data capacity %" PRId64 " < required %" PRId64,
                    out->data_capacity, bytes + (last - first));
      memcpy(out->data + bytes, c.data + first, (size_t)(last - first));
      bytes += last - first;
    }
    if (materialized || keep) {
      for (int64_t k = 0; k < len; ++k)
        if (!c.array.validity || get_bit(c.array.validity, c.array.validity_offset + k)) set_bit(out->array.validity, row + k);
    }
    row += len;
  }
  out->data_len = bytes;
  if (materialized || keep) {
    out->array.has_validity = 1;
    out->array.null_count = total - count_bits(out->array.validity, 0, total);
  }
  return ACU_OK;
}


// ---- Utf8View / BinaryView buffer management of BatchCoalescer (arrow-select/src/coalesce/byte_view.rs) -------------------
// A view = 16 bytes: length u32 | 12 inline bytes, or length | 4-byte prefix | buffer index u32 | offset u32.
static inline uint32_t view_len(const uint8_t *views, int64_t i) {
  uint32_t l;
  memcpy(&l, views + 16 * i, 4);
  return l;
}

// GenericByteViewArray::total_buffer_bytes_used (arrow-array/src/array/byte_view_array.rs:749-761)
int64_t orc_view_bytes_used(const uint8_t *views, int64_t n) {
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t len = view_len(views, i);
    if (len > 12) total += len;
  }
  return total;
}

// the loop of append_views_and_copy_strings that decides how many views still go to the current buffer
// (coalesce/byte_view.rs:259-271): every view's length is compared with what is left, long views consume it
void orc_view_fit(const uint8_t *views, int64_t n, int64_t remaining_capacity, int64_t *out_views, int64_t *out_bytes) {
  int64_t remaining = remaining_capacity, num = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t str_len = view_len(views, i);
    if (remaining < str_len) break;
    if (str_len > 12) remaining -= str_len;
    ++num;
  }
  *out_views = num;
  *out_bytes = remaining_capacity - remaining;
}

// append_views_and_copy_strings_inner (coalesce/byte_view.rs:298-354): returns the bytes appended to dst
int64_t orc_view_copy_strings(const uint8_t *views, int64_t n, const uint8_t *const *buffers, uint32_t new_buffer_index, uint8_t *dst,
                              int64_t dst_len, uint8_t *out_views) {
  int64_t pos = dst_len;
  for (int64_t i = 0; i < n; ++i) {
    uint8_t v[16];
    memcpy(v, views + 16 * i, 16);
    const uint32_t len = view_len(views, i);
    if (len > 12) {
      uint32_t buffer_index, offset;
      memcpy(&buffer_index, v + 8, 4);
      memcpy(&offset, v + 12, 4);
      const uint32_t new_offset = (uint32_t)pos;  // b.offset = dst_buffer.len() as u32
      memcpy(v + 8, &new_buffer_index, 4);
      memcpy(v + 12, &new_offset, 4);
      memcpy(dst + pos, buffers[buffer_index] + offset, len);
      pos += len;
    }
    memcpy(out_views + 16 * i, v, 16);
  }
  return pos - dst_len;
}

// append_views_and_update_buffer_index (coalesce/byte_view.rs:176-216)
void orc_view_rebase(const uint8_t *views, int64_t n, uint32_t delta, uint8_t *out_views) {
  for (int64_t i = 0; i < n; ++i) {
    uint8_t v[16];
    memcpy(v, views + 16 * i, 16);
    if (view_len(views, i) > 12) {
      uint32_t buffer_index;
      memcpy(&buffer_index, v + 8, 4);
      buffer_index += delta;
      memcpy(v + 8, &buffer_index, 4);
    }
    memcpy(out_views + 16 * i, v, 16);
  }
}


// SlicesIterator over prep_null_mask_filter(predicate) (arrow-select/src/filter.rs:44-77,167-171; BitSliceIterator,
// arrow-buffer/src/util/bit_iterator.rs): [start, end) of every maximal run of rows with value && valid. Returns the run count;
// pairs beyond `capacity` are not written.
int64_t orc_filter_slices(const acu_array *pred, uint64_t *out_pairs, int64_t capacity) {
  const uint8_t *vals = static_cast<const uint8_t *>(pred->values);
  int64_t n_slices = 0, start = -1;
  for (int64_t i = 0; i < pred->len; ++i) {
    bool set = get_bit(vals, pred->values_offset + i);
    if (set && pred->validity) set = get_bit(pred->validity, pred->validity_offset + i);
    if (set && start < 0) start = i;
    if (!set && start >= 0) {
      if (n_slices < capacity) { out_pairs[2 * n_slices] = (uint64_t)start; out_pairs[2 * n_slices + 1] = (uint64_t)i; }
      ++n_slices;
      start = -1;
    }
  }
  if (start >= 0) {
    if (n_slices < capacity) { out_pairs[2 * n_slices] = (uint64_t)start; out_pairs[2 * n_slices + 1] = (uint64_t)pred->len; }
    ++n_slices;
  }
  return n_slices;
}

}  // extern "C"
