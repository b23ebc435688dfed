// boolean.cu — predicate construction on device: the bitmap kernels that sit immediately
// before filter, so that a `cmp -> and -> filter` pipeline never
// round-trips its mask through the host.
//
//   and / or / and_not      arrow-arith/src/boolean.rs:256-303 (binary_boolean_kernel :224-243:
//                           values = op over ALL slots, nulls = NullBuffer::union)
//   and_kleene / or_kleene  boolean.rs:60-124, :156-222 (three-valued logic; the validity formula
//                           depends on which side carries a NullBuffer)
//   not                     boolean.rs:310-314
//   is_null / is_not_null   boolean.rs:327-354 (no NullBuffer on the result)
//
// One thread per output u64 word (64 rows): up to four input words with arbitrary bit offsets,
// two output words, popcount of the validity into the result block. HBM-bound on
// len/8 x (inputs + outputs) bytes.
#include "bitmap.cuh"

namespace {

struct BoolArgs {
  const uint8_t *av;  // a values bitmap (NULL for is_null / is_not_null)
  int64_t aoff;
  const uint8_t *an;  // a validity or NULL
  int64_t anoff;
  const uint8_t *bv;  // b values (binary ops)
  int64_t boff;
  const uint8_t *bn;
  int64_t bnoff;
  int64_t len;
  uint64_t *out_v;
  uint64_t *out_n;    // NULL: the result has no NullBuffer
  unsigned long long *res;
};

template <int OP>
__device__ __forceinline__ void bool_word(bool has_an, bool has_bn, uint64_t A, uint64_t B, uint64_t AN, uint64_t BN, uint64_t lenmask, uint64_t *v, uint64_t *n) {
  *n = lenmask;
  if (OP == ACU_BOOL_AND) { *v = A & B; *n = AN & BN; }
  else if (OP == ACU_BOOL_OR) { *v = A | B; *n = AN & BN; }
  else if (OP == ACU_BOOL_AND_NOT) { *v = A & ~B; *n = AN & BN; }
  else if (OP == ACU_BOOL_AND_KLEENE) {
    *v = A & B;
    if (has_an && has_bn) *n = (AN | (BN & ~B)) & (BN | (AN & ~A));  // boolean.rs:99-118
    else if (has_an) *n = AN | ~B;                                   // boolean.rs:71-84
    else if (has_bn) *n = BN | ~A;                                   // boolean.rs:85-95
  } else if (OP == ACU_BOOL_OR_KLEENE) {
    *v = A | B;
    if (has_an && has_bn) *n = (AN | (BN & B)) & (BN | (AN & A));    // boolean.rs:195-214
    else if (has_an) *n = AN | B;
    else if (has_bn) *n = BN | A;
  } else if (OP == ACU_BOOL_NOT) { *v = ~A; *n = AN; }
  else if (OP == ACU_BOOL_IS_NULL) { *v = ~AN; }
  else { *v = AN; }
  *v &= lenmask;
  *n &= lenmask;
}

// ALIGNED: every input bitmap starts on a u64 word (8-byte aligned base, bit offset a multiple of 64 — the
// common case: unsliced arrays and slices on 64-row boundaries): plain coalesced u64 loads, no funnel shifts.
template <int OP, bool ALIGNED>
__global__ void __launch_bounds__(256) k_boolean(const BoolArgs a) {
  const int64_t words = (a.len + 63) >> 6;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint64_t *av = reinterpret_cast<const uint64_t *>(a.av) + (a.aoff >> 6), *bv = reinterpret_cast<const uint64_t *>(a.bv) + (a.boff >> 6);
  const uint64_t *an = reinterpret_cast<const uint64_t *>(a.an) + (a.anoff >> 6), *bn = reinterpret_cast<const uint64_t *>(a.bn) + (a.bnoff >> 6);
  unsigned cnt = 0;
#pragma unroll 2
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    const int64_t pos = w << 6;
    const int64_t left = a.len - pos;
    const uint64_t lenmask = left >= 64 ? ~0ull : ((1ull << left) - 1ull);
    uint64_t A, B, AN, BN;
    if (ALIGNED) {
      A = a.av ? __ldg(av + w) : 0ull;
      B = a.bv ? __ldg(bv + w) : 0ull;
      AN = a.an ? __ldg(an + w) : lenmask;
      BN = a.bn ? __ldg(bn + w) : lenmask;
    } else {
      A = a.av ? ld_bits64(a.av, a.aoff + pos, a.aoff + a.len) : 0ull;
      B = a.bv ? ld_bits64(a.bv, a.boff + pos, a.boff + a.len) : 0ull;
      AN = a.an ? ld_bits64(a.an, a.anoff + pos, a.anoff + a.len) : lenmask;
      BN = a.bn ? ld_bits64(a.bn, a.bnoff + pos, a.bnoff + a.len) : lenmask;
    }
    uint64_t v, n;
    bool_word<OP>(a.an != nullptr, a.bn != nullptr, A, B, AN, BN, lenmask, &v, &n);
    a.out_v[w] = v;
    if (a.out_n) {
      a.out_n[w] = n;
      cnt += __popcll(n);
    }
  }
  if (a.out_n) {
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)cnt);
  }
}

template <int OP>
acu_status launch(acu_ctx *ctx, const BoolArgs &a) {
  const int64_t words = (a.len + 63) >> 6;
  auto word_aligned = [](const uint8_t *p, int64_t off) { return p == nullptr || (((uintptr_t)p & 7) == 0 && (off & 63) == 0); };
  // (the aligned loads read whole words: the last one may extend past len inside the allocation's 64-bit padding)
  const bool aligned = word_aligned(a.av, a.aoff) && word_aligned(a.an, a.anoff) && word_aligned(a.bv, a.boff) && word_aligned(a.bn, a.bnoff);
  if (aligned) ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_boolean<OP, true>), acu_grid(ctx, (words + 255) / 256, 16), 256, 0, a);
  else ACU_LAUNCH_TIMED(ctx, ACU_K_CMP, (k_boolean<OP, false>), acu_grid(ctx, (words + 255) / 256, 16), 256, 0, a);
  return ACU_OK;
}

}  // namespace

extern "C" acu_status acu_boolean(acu_ctx *ctx, acu_bool_op op, const acu_array *a, const acu_array *b, acu_array_out *out) {
  ACU_ENTER(ctx);
  const bool binary = op <= ACU_BOOL_OR_KLEENE;
  if ((int)op < 0 || op > ACU_BOOL_IS_NOT_NULL) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "boolean: unknown op %d", (int)op);
  if (binary && b == nullptr) return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "boolean: binary op needs two arrays");
  if (binary && a->len != b->len)  // boolean.rs:61-65, :232-236
    return acu_fail(ctx, ACU_ERR_COMPUTE, -1, 0, 0, 0, "Cannot perform bitwise operation on arrays of different length");
  const int64_t n = a->len;
  out->len = n;
  out->null_count = 0;
  const bool nulls_out = op == ACU_BOOL_NOT ? a->validity != nullptr
                         : binary            ? (a->validity != nullptr || b->validity != nullptr)
                                             : false;
  out->has_validity = nulls_out ? 1 : 0;
  if (n == 0) return ACU_OK;
  BoolArgs ba{};
  const bool values_in = op != ACU_BOOL_IS_NULL && op != ACU_BOOL_IS_NOT_NULL;
  ba.av = values_in ? static_cast<const uint8_t *>(a->values) : nullptr;
  ba.aoff = a->values_offset;
  ba.an = a->validity;
  ba.anoff = a->validity_offset;
  if (binary) {
    ba.bv = static_cast<const uint8_t *>(b->values);
    ba.boff = b->values_offset;
    ba.bn = b->validity;
    ba.bnoff = b->validity_offset;
  }
  ba.len = n;
  ba.out_v = static_cast<uint64_t *>(out->values);
  ba.out_n = nulls_out ? reinterpret_cast<uint64_t *>(out->validity) : nullptr;
  ba.res = acu_dres(ctx, 0);
  if (nulls_out) ACU_TRY(acu_res_reset(ctx));
  switch (op) {
    case ACU_BOOL_AND: ACU_TRY(launch<ACU_BOOL_AND>(ctx, ba)); break;
    case ACU_BOOL_OR: ACU_TRY(launch<ACU_BOOL_OR>(ctx, ba)); break;
    case ACU_BOOL_AND_NOT: ACU_TRY(launch<ACU_BOOL_AND_NOT>(ctx, ba)); break;
    case ACU_BOOL_AND_KLEENE: ACU_TRY(launch<ACU_BOOL_AND_KLEENE>(ctx, ba)); break;
    case ACU_BOOL_OR_KLEENE: ACU_TRY(launch<ACU_BOOL_OR_KLEENE>(ctx, ba)); break;
    case ACU_BOOL_NOT: ACU_TRY(launch<ACU_BOOL_NOT>(ctx, ba)); break;
    case ACU_BOOL_IS_NULL: ACU_TRY(launch<ACU_BOOL_IS_NULL>(ctx, ba)); break;
    default: ACU_TRY(launch<ACU_BOOL_IS_NOT_NULL>(ctx, ba)); break;
  }
  if (nulls_out) {
    ACU_TRY(acu_res_fetch(ctx));
    out->null_count = n - (int64_t)ctx->h_res[RES_COUNT];
  } else {
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    acu_kstats_drain(ctx);
  }
  return ACU_OK;
}
