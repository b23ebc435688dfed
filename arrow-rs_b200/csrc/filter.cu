// filter.cu — arrow-select/src/filter.rs on the device.
//
//   FilterBuilder::new/optimize/build  (filter.rs:254-324)  -> acu_filter_plan_create
//   FilterPredicate::filter            (filter.rs:449-452)  -> acu_filter_primitive / _boolean
//   filter_native / filter_bits / filter_nulls (filter.rs:512-533, :680-788)
//
// Design (stream compaction, HBM-bound):
//   plan:    one pass over the predicate bits (N/8 bytes): mask = values & validity
//            normalised to bit offset 0, per-tile (4096 rows) popcounts, two-level exclusive
//            scan -> every tile knows its first output row. The plan is reused by every
//            column of a RecordBatch (FilterPredicate::filter_record_batch, filter.rs:459-478).
//   compact: one CTA per tile, grid-stride. 128-bit loads of the tile's values are
//            PREDICATED on "this 16-byte chunk holds a selected row", so at low selectivity
//            whole 32-B DRAM sectors are never fetched; selected elements are ranked with
//            popcounts of the mask word prefix, staged in shared memory and written out as
//            full coalesced lines. The validity bits of the selected rows are compacted in the
//            same pass (filter_bits) with their popcount (filter_nulls).
#include "bitmap.cuh"
#include "internal.cuh"

#define TILE_ROWS 4096
#define TILE_WORDS (TILE_ROWS / 64)
#define SCAN_CHUNK 4096  // tiles per scan block

struct acu_filter_plan {
  int64_t len = 0;
  int64_t count = 0;
  int32_t strategy = ACU_FILTER_NONE;
  int64_t n_tiles = 0;
  uint64_t *mask = nullptr;         // n_tiles * TILE_WORDS words, zero padded
  uint32_t *tile_local = nullptr;   // exclusive prefix of tile counts inside its scan chunk
  uint32_t *tile_count = nullptr;   // selected rows per tile
  uint64_t *chunk_offset = nullptr; // exclusive prefix of chunk totals
  void *storage = nullptr;
};

namespace {

// ---- plan kernels ----------------------------------------------------------------------
// One warp per tile: lane l owns mask words l and l+32 of the tile.
__global__ void __launch_bounds__(256) k_plan_mask(const uint8_t *__restrict__ pv, int64_t poff,
                                                   const uint8_t *__restrict__ nv, int64_t noff,
                                                   int64_t len, int64_t n_tiles, uint64_t *__restrict__ mask,
                                                   uint32_t *__restrict__ tile_count) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tiles; t += nwarps) {
    unsigned c = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t w = t * TILE_WORDS + h * 32 + lane;
      const int64_t row = w << 6;
      uint64_t m = ld_bits64(pv, poff + row, poff + len);
      if (nv) m &= ld_bits64(nv, noff + row, noff + len);  // prep_null_mask_filter
      mask[w] = m;
      c += __popcll(m);
    }
    c = warp_sum(c);
    if (lane == 0) tile_count[t] = c;
  }
}

// Block-wide exclusive scan of up to SCAN_CHUNK tile counts (1024 threads x 4).
__global__ void __launch_bounds__(1024) k_plan_scan_chunks(const uint32_t *__restrict__ tile_count, int64_t n_tiles,
                                                           uint32_t *__restrict__ tile_local,
                                                           uint64_t *__restrict__ chunk_total) {
  __shared__ uint32_t warp_tot[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * 4;
  uint32_t c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k] = (base + k < n_tiles) ? tile_count[base + k] : 0u;
  uint32_t mine = c[0] + c[1] + c[2] + c[3];
  uint32_t incl = mine;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(ACU_FULL_MASK, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(ACU_FULL_MASK, wi, o);
      if (lane >= o) wi += y;
    }
    warp_tot[lane] = wi - w;  // exclusive
    if (lane == 31) chunk_total[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t excl = warp_tot[wid] + incl - mine;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n_tiles) tile_local[base + k] = excl;
    excl += c[k];
  }
}

// Single block: exclusive scan of the chunk totals (u64), grand total -> res[RES_COUNT].
__global__ void __launch_bounds__(1024) k_plan_scan_top(uint64_t *__restrict__ chunk_total, int64_t n_chunks,
                                                        unsigned long long *__restrict__ res) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_chunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    uint64_t v = i < n_chunks ? chunk_total[i] : 0ull, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t y = __shfl_up_sync(ACU_FULL_MASK, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      uint64_t w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(ACU_FULL_MASK, wi, o);
        if (lane >= o) wi += y;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const uint64_t carry = carry_s;
    if (i < n_chunks) chunk_total[i] = carry + warp_tot[wid] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + warp_tot[31] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) res[RES_COUNT] = carry_s;
}

// ---- compaction kernel ------------------------------------------------------------------
template <int W> struct ElemOf;
template <> struct ElemOf<1> { using type = uint8_t; };
template <> struct ElemOf<2> { using type = uint16_t; };
template <> struct ElemOf<4> { using type = uint32_t; };
template <> struct ElemOf<8> { using type = uint64_t; };

struct FilterArgs {
  const uint8_t *values;
  uint8_t *out;
  const uint64_t *mask;
  const uint32_t *tile_local;
  const uint32_t *tile_count;
  const uint64_t *chunk_offset;
  int64_t n_tiles;
  const uint8_t *validity;  // NULL: no validity compaction
  int64_t voff;
  uint32_t *out_valid;      // pre-zeroed, u32 words
  unsigned long long *res;
  int aligned16;
};

// W <= 8: selected elements staged through shared memory, coalesced stores.
template <int W, bool HAS_VALID>
__global__ void __launch_bounds__(256) k_filter_small(const FilterArgs a) {
  using E = typename ElemOf<W>::type;
  constexpr int RPC = 16 / W;              // rows per 16-byte chunk
  constexpr int ROUNDS = TILE_ROWS / (256 * RPC);  // chunks per thread (= W)
  constexpr int BATCH = ROUNDS < 8 ? ROUNDS : 8;
  __shared__ uint64_t s_mask[TILE_WORDS];
  __shared__ uint16_t s_pref[TILE_WORDS];
  __shared__ __align__(16) E s_vals[TILE_ROWS];
  __shared__ uint8_t s_vbit[HAS_VALID ? TILE_ROWS : 4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned valid_cnt = 0;

  for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint32_t cnt = a.tile_count[t];
    if (cnt == 0) continue;  // uniform per CTA
    const int64_t row0 = t * TILE_ROWS;
    const uint64_t out0 = a.chunk_offset[t / SCAN_CHUNK] + a.tile_local[t];
    __syncthreads();  // previous tile's readers of s_vals / s_vbit are done
    // ---- phase A: mask words + per-word exclusive prefix (one warp, 2 words per lane) ----
    if (wid == 0) {
      const uint64_t m0 = a.mask[t * TILE_WORDS + 2 * lane], m1 = a.mask[t * TILE_WORDS + 2 * lane + 1];
      const uint32_t c0 = __popcll(m0), c1 = __popcll(m1);
      uint32_t incl = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(ACU_FULL_MASK, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t excl = incl - (c0 + c1);
      s_mask[2 * lane] = m0;
      s_mask[2 * lane + 1] = m1;
      s_pref[2 * lane] = (uint16_t)excl;
      s_pref[2 * lane + 1] = (uint16_t)(excl + c0);
    }
    __syncthreads();
    // ---- phase B: predicated loads, rank, scatter into shared memory ----
    const uint8_t *tile_src = a.values + (size_t)row0 * W;
#pragma unroll
    for (int b0 = 0; b0 < ROUNDS; b0 += BATCH) {
      uint4 v[BATCH];
      uint32_t bits[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        const int c = (b0 + j) * 256 + tid;  // chunk index inside the tile
        const int r = c * RPC;               // first row of the chunk
        bits[j] = (uint32_t)(s_mask[r >> 6] >> (r & 63)) & ((1u << RPC) - 1u);
        v[j] = make_uint4(0, 0, 0, 0);
        if (bits[j]) {
          if (a.aligned16) {
            v[j] = ld_stream16(tile_src + (size_t)c * 16);
          } else {  // sliced array whose base is not 16-B aligned: element-wise loads
            E *ve = reinterpret_cast<E *>(&v[j]);
#pragma unroll
            for (int e = 0; e < RPC; ++e)
              if ((bits[j] >> e) & 1u) ve[e] = __ldg(reinterpret_cast<const E *>(tile_src) + r + e);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        if (!bits[j]) continue;
        const int c = (b0 + j) * 256 + tid;
        const int r = c * RPC;
        const uint64_t word = s_mask[r >> 6];
        uint32_t rank = s_pref[r >> 6] + __popcll(word & ((1ull << (r & 63)) - 1ull));
        const E *ve = reinterpret_cast<const E *>(&v[j]);
#pragma unroll
        for (int e = 0; e < RPC; ++e) {
          if ((bits[j] >> e) & 1u) {
            s_vals[rank] = ve[e];
            if (HAS_VALID) s_vbit[rank] = (uint8_t)ld_bit(a.validity, a.voff + row0 + r + e);
            ++rank;
          }
        }
      }
    }
    __syncthreads();
    // ---- phase C: coalesced copy-out + validity bit packing ----
    E *dst = reinterpret_cast<E *>(a.out) + out0;
    for (uint32_t k = tid; k < cnt; k += 256) dst[k] = s_vals[k];
    if (HAS_VALID) {
      const uint64_t first_w = out0 >> 5, last_w = (out0 + cnt - 1) >> 5;
      for (uint64_t w = first_w + wid; w <= last_w; w += 8) {
        const int64_t k = (int64_t)(w << 5) + lane - (int64_t)out0;
        const bool bit = (k >= 0 && k < (int64_t)cnt) ? (s_vbit[k] != 0) : false;
        const uint32_t word = __ballot_sync(ACU_FULL_MASK, bit);
        if (lane == 0) {
          if (w == first_w || w == last_w) atomicOr(a.out_valid + w, word);  // shared with neighbours
          else a.out_valid[w] = word;
          valid_cnt += __popc(word);
        }
      }
    }
  }
  if (HAS_VALID && lane == 0 && valid_cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)valid_cnt);
}

// W = 16 / 32 (Decimal128/256, IntervalMonthDayNano): rows are whole 16-byte chunks, written
// straight to their final position.
template <int W, bool HAS_VALID>
__global__ void __launch_bounds__(256) k_filter_wide(const FilterArgs a) {
  constexpr int CPR = W / 16;  // chunks per row
  constexpr int ROUNDS = TILE_ROWS * CPR / 256;
  __shared__ uint64_t s_mask[TILE_WORDS];
  __shared__ uint16_t s_pref[TILE_WORDS];
  __shared__ uint8_t s_vbit[HAS_VALID ? TILE_ROWS : 4];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned valid_cnt = 0;
  for (int64_t t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const uint32_t cnt = a.tile_count[t];
    if (cnt == 0) continue;
    const int64_t row0 = t * TILE_ROWS;
    const uint64_t out0 = a.chunk_offset[t / SCAN_CHUNK] + a.tile_local[t];
    __syncthreads();
    if (wid == 0) {
      const uint64_t m0 = a.mask[t * TILE_WORDS + 2 * lane], m1 = a.mask[t * TILE_WORDS + 2 * lane + 1];
      const uint32_t c0 = __popcll(m0), c1 = __popcll(m1);
      uint32_t incl = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(ACU_FULL_MASK, incl, o);
        if (lane >= o) incl += y;
      }
      const uint32_t excl = incl - (c0 + c1);
      s_mask[2 * lane] = m0;
      s_mask[2 * lane + 1] = m1;
      s_pref[2 * lane] = (uint16_t)excl;
      s_pref[2 * lane + 1] = (uint16_t)(excl + c0);
    }
    __syncthreads();
    const uint8_t *tile_src = a.values + (size_t)row0 * W;
    for (int b0 = 0; b0 < ROUNDS; b0 += 8) {
      uint4 v[8];
      bool sel[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (b0 + j) * 256 + tid;
        const int r = c / CPR;
        sel[j] = (s_mask[r >> 6] >> (r & 63)) & 1ull;
        if (sel[j]) {
          if (a.aligned16) v[j] = ld_stream16(tile_src + (size_t)c * 16);
          else {
            const uint64_t *p = reinterpret_cast<const uint64_t *>(tile_src + (size_t)c * 16);
            uint64_t lo = __ldg(p), hi = __ldg(p + 1);
            v[j] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!sel[j]) continue;
        const int c = (b0 + j) * 256 + tid;
        const int r = c / CPR, half = c % CPR;
        const uint32_t rank = s_pref[r >> 6] + __popcll(s_mask[r >> 6] & ((1ull << (r & 63)) - 1ull));
        uint64_t *q = reinterpret_cast<uint64_t *>(a.out + (out0 + rank) * W + half * 16);  // 8-B aligned at least
        q[0] = (uint64_t)v[j].x | ((uint64_t)v[j].y << 32);
        q[1] = (uint64_t)v[j].z | ((uint64_t)v[j].w << 32);
        if (HAS_VALID && half == 0) s_vbit[rank] = (uint8_t)ld_bit(a.validity, a.voff + row0 + r);
      }
    }
    if (HAS_VALID) {
      __syncthreads();
      const uint64_t first_w = out0 >> 5, last_w = (out0 + cnt - 1) >> 5;
      for (uint64_t w = first_w + wid; w <= last_w; w += 8) {
        const int64_t k = (int64_t)(w << 5) + lane - (int64_t)out0;
        const bool bit = (k >= 0 && k < (int64_t)cnt) ? (s_vbit[k] != 0) : false;
        const uint32_t word = __ballot_sync(ACU_FULL_MASK, bit);
        if (lane == 0) {
          if (w == first_w || w == last_w) atomicOr(a.out_valid + w, word);
          else a.out_valid[w] = word;
          valid_cnt += __popc(word);
        }
      }
    }
  }
  if (HAS_VALID && lane == 0 && valid_cnt) atomicAdd(a.res + RES_COUNT, (unsigned long long)valid_cnt);
}

// Bit compaction for boolean VALUES (filter_boolean -> filter_bits): same plan, 1 bit/row.
// One warp per tile; each lane ranks the set bits of two mask words and scatters the
// selected source bits with atomicOr into the pre-zeroed output.
__global__ void __launch_bounds__(256) k_filter_bits(const uint8_t *__restrict__ src, int64_t soff,
                                                     const uint64_t *__restrict__ mask,
                                                     const uint32_t *__restrict__ tile_local,
                                                     const uint32_t *__restrict__ tile_count,
                                                     const uint64_t *__restrict__ chunk_offset, int64_t n_tiles,
                                                     int64_t len, uint32_t *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp; t < n_tiles; t += nwarps) {
    if (tile_count[t] == 0) continue;
    const uint64_t out0 = chunk_offset[t / SCAN_CHUNK] + tile_local[t];
    const uint64_t m0 = mask[t * TILE_WORDS + 2 * lane], m1 = mask[t * TILE_WORDS + 2 * lane + 1];
    const uint32_t c0 = __popcll(m0), c1 = __popcll(m1);
    uint32_t incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(ACU_FULL_MASK, incl, o);
      if (lane >= o) incl += y;
    }
    uint64_t k = out0 + incl - (c0 + c1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint64_t m = h ? m1 : m0;
      const int64_t row = (t * TILE_WORDS + 2 * lane + h) << 6;
      const uint64_t sb = ld_bits64(src, soff + row, soff + len);
      while (m) {
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1;
        if ((sb >> b) & 1ull) atomicOr(out + (k >> 5), 1u << (k & 31));
        ++k;
      }
    }
  }
}

acu_status check_len(acu_ctx *ctx, const acu_filter_plan *plan, int64_t values_len) {
  if (plan->len > values_len)  // filter.rs:536-542
    return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, (uint64_t)values_len,
                    "Filter predicate of length %lld is larger than target array of length %lld",
                    (long long)plan->len, (long long)values_len);
  return ACU_OK;
}

// `values.slice(0, count)` nulls for IterationStrategy::All (filter.rs:546)
acu_status slice_nulls(acu_ctx *ctx, const acu_array *a, int64_t count, acu_array_out *out) {
  out->has_validity = 0;
  out->null_count = 0;
  if (!a->validity || count == 0) { out->has_validity = a->validity != nullptr; return ACU_OK; }
  ACU_TRY(acu_res_reset(ctx));
  ACU_TRY(acu_bitmap_and_launch(ctx, a->validity, a->validity_offset, nullptr, 0, count,
                                reinterpret_cast<uint64_t *>(out->validity), true));
  ACU_TRY(acu_res_fetch(ctx));
  out->has_validity = 1;
  out->null_count = count - (int64_t)ctx->h_res[RES_COUNT];
  return ACU_OK;
}

template <int W>
acu_status launch_filter(acu_ctx *ctx, const FilterArgs &fa, bool has_valid) {
  if constexpr (W <= 8) {
    if (has_valid) ACU_LAUNCH(ctx, (k_filter_small<W, true>), acu_wave_grid(ctx, k_filter_small<W, true>, 256, 0, fa.n_tiles), 256, 0, fa);
    else ACU_LAUNCH(ctx, (k_filter_small<W, false>), acu_wave_grid(ctx, k_filter_small<W, false>, 256, 0, fa.n_tiles), 256, 0, fa);
  } else {
    if (has_valid) ACU_LAUNCH(ctx, (k_filter_wide<W, true>), acu_wave_grid(ctx, k_filter_wide<W, true>, 256, 0, fa.n_tiles), 256, 0, fa);
    else ACU_LAUNCH(ctx, (k_filter_wide<W, false>), acu_wave_grid(ctx, k_filter_wide<W, false>, 256, 0, fa.n_tiles), 256, 0, fa);
  }
  return ACU_OK;
}

}  // namespace

extern "C" {

acu_status acu_filter_plan_create(acu_ctx *ctx, const acu_array *pred, acu_filter_plan **out_plan) {
  *out_plan = nullptr;
  acu_filter_plan *plan = new acu_filter_plan();
  const int64_t len = pred->len;
  plan->len = len;
  if (len == 0) { *out_plan = plan; return ACU_OK; }
  acu_status st;
  const int64_t nc = acu_resolve_null_count(ctx, pred, &st);
  if (st != ACU_OK) { delete plan; return st; }
  const int64_t n_tiles = (len + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t n_chunks = (n_tiles + SCAN_CHUNK - 1) / SCAN_CHUNK;
  plan->n_tiles = n_tiles;
  size_t mask_b = (size_t)n_tiles * TILE_WORDS * 8, tl_b = ((size_t)n_tiles * 4 + 255) & ~(size_t)255,
         co_b = ((size_t)n_chunks * 8 + 255) & ~(size_t)255;
  void *mem = nullptr;
  st = acu_malloc(ctx, mask_b + 2 * tl_b + co_b, &mem);
  if (st != ACU_OK) { delete plan; return st; }
  plan->storage = mem;
  plan->mask = static_cast<uint64_t *>(mem);
  plan->tile_local = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(mem) + mask_b);
  plan->tile_count = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(mem) + mask_b + tl_b);
  plan->chunk_offset = reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(mem) + mask_b + 2 * tl_b);
  auto bail = [&](acu_status s) { acu_free(ctx, mem); delete plan; return s; };
  const uint8_t *nv = (pred->validity && nc > 0) ? pred->validity : nullptr;  // filter.rs:261-264
  {
    acu_status s = acu_res_reset(ctx);
    if (s != ACU_OK) return bail(s);
  }
  k_plan_mask<<<acu_grid(ctx, (n_tiles + 7) / 8, 8), 256, 0, ctx->stream>>>(
      static_cast<const uint8_t *>(pred->values), pred->values_offset, nv, pred->validity_offset, len, n_tiles,
      plan->mask, plan->tile_count);
  k_plan_scan_chunks<<<(unsigned)n_chunks, 1024, 0, ctx->stream>>>(plan->tile_count, n_tiles, plan->tile_local, plan->chunk_offset);
  k_plan_scan_top<<<1, 1024, 0, ctx->stream>>>(plan->chunk_offset, n_chunks, ctx->d_res);
  ctx->launches += 3;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return bail(acu_cuda_fail(ctx, e, "filter plan kernels"));
  {
    acu_status s = acu_res_fetch(ctx);
    if (s != ACU_OK) return bail(s);
  }
  plan->count = (int64_t)ctx->h_res[RES_COUNT];
  // IterationStrategy::default_strategy (filter.rs:346-364)
  if (plan->count == 0) plan->strategy = ACU_FILTER_NONE;
  else if (plan->count == len) plan->strategy = ACU_FILTER_ALL;
  else if ((double)plan->count / (double)len > 0.8) plan->strategy = ACU_FILTER_SLICES;
  else plan->strategy = ACU_FILTER_INDEX;
  *out_plan = plan;
  return ACU_OK;
}

void acu_filter_plan_destroy(acu_ctx *ctx, acu_filter_plan *plan) {
  if (!plan) return;
  if (plan->storage) acu_free(ctx, plan->storage);
  delete plan;
}
int64_t acu_filter_plan_count(const acu_filter_plan *plan) { return plan->count; }
int64_t acu_filter_plan_len(const acu_filter_plan *plan) { return plan->len; }
int32_t acu_filter_plan_strategy(const acu_filter_plan *plan) { return plan->strategy; }

acu_status acu_filter_primitive(acu_ctx *ctx, const acu_filter_plan *plan, int32_t elem_bytes,
                                const acu_array *values, acu_array_out *out) {
  ACU_TRY(check_len(ctx, plan, values->len));
  out->len = plan->count;
  out->has_validity = 0;
  out->null_count = 0;
  if (plan->strategy == ACU_FILTER_NONE) return ACU_OK;
  if (plan->strategy == ACU_FILTER_ALL) {
    ACU_CUDA(ctx, cudaMemcpyAsync(out->values, values->values, (size_t)plan->count * elem_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ACU_TRY(slice_nulls(ctx, values, plan->count, out));
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACU_OK;
  }
  acu_status st;
  const int64_t nc = acu_resolve_null_count(ctx, values, &st);
  ACU_TRY(st);
  const bool has_valid = values->validity && nc > 0;  // filter_nulls (filter.rs:512-516)
  FilterArgs fa{};
  fa.values = static_cast<const uint8_t *>(values->values);
  fa.out = static_cast<uint8_t *>(out->values);
  fa.mask = plan->mask;
  fa.tile_local = plan->tile_local;
  fa.tile_count = plan->tile_count;
  fa.chunk_offset = plan->chunk_offset;
  fa.n_tiles = plan->n_tiles;
  fa.validity = has_valid ? values->validity : nullptr;
  fa.voff = values->validity_offset;
  fa.out_valid = reinterpret_cast<uint32_t *>(out->validity);
  fa.res = ctx->d_res;
  fa.aligned16 = ((uintptr_t)values->values % 16) == 0;
  if (has_valid) {
    ACU_TRY(acu_res_reset(ctx));
    ACU_CUDA(ctx, cudaMemsetAsync(out->validity, 0, acu_bitmap_bytes(plan->count), ctx->stream));
  }
  switch (elem_bytes) {
    case 1: ACU_TRY(launch_filter<1>(ctx, fa, has_valid)); break;
    case 2: ACU_TRY(launch_filter<2>(ctx, fa, has_valid)); break;
    case 4: ACU_TRY(launch_filter<4>(ctx, fa, has_valid)); break;
    case 8: ACU_TRY(launch_filter<8>(ctx, fa, has_valid)); break;
    case 16: ACU_TRY(launch_filter<16>(ctx, fa, has_valid)); break;
    case 32: ACU_TRY(launch_filter<32>(ctx, fa, has_valid)); break;
    default:
      return acu_fail(ctx, ACU_ERR_INVALID_ARGUMENT, -1, 0, 0, 0, "filter: unsupported element width %d", elem_bytes);
  }
  if (has_valid) {
    ACU_TRY(acu_res_fetch(ctx));
    const int64_t null_count = plan->count - (int64_t)ctx->h_res[RES_COUNT];
    if (null_count > 0) {  // filter.rs:523-525: None when the filtered result has no nulls
      out->has_validity = 1;
      out->null_count = null_count;
    }
  } else {
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return ACU_OK;
}

acu_status acu_filter_boolean(acu_ctx *ctx, const acu_filter_plan *plan, const acu_array *values,
                              acu_array_out *out) {
  ACU_TRY(check_len(ctx, plan, values->len));
  out->len = plan->count;
  out->has_validity = 0;
  out->null_count = 0;
  if (plan->strategy == ACU_FILTER_NONE) return ACU_OK;
  if (plan->strategy == ACU_FILTER_ALL) {
    ACU_TRY(acu_bitmap_and_launch(ctx, static_cast<const uint8_t *>(values->values), values->values_offset, nullptr, 0,
                                  plan->count, static_cast<uint64_t *>(out->values), false));
    ACU_TRY(slice_nulls(ctx, values, plan->count, out));
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ACU_OK;
  }
  const int grid = acu_grid(ctx, (plan->n_tiles + 7) / 8, 8);
  ACU_CUDA(ctx, cudaMemsetAsync(out->values, 0, acu_bitmap_bytes(plan->count), ctx->stream));
  ACU_LAUNCH(ctx, k_filter_bits, grid, 256, 0, static_cast<const uint8_t *>(values->values), values->values_offset,
             plan->mask, plan->tile_local, plan->tile_count, plan->chunk_offset, plan->n_tiles, plan->len,
             static_cast<uint32_t *>(out->values));
  return acu_filter_nulls_internal(ctx, plan, values, out);
}

}  // extern "C"

// FilterPredicate::filter_nulls (filter.rs:512-533): bit-compact the validity, count, drop
// the buffer when the result has no nulls. Synchronises the stream.
acu_status acu_filter_nulls_internal(acu_ctx *ctx, const acu_filter_plan *plan, const acu_array *a,
                                     acu_array_out *out) {
  out->has_validity = 0;
  out->null_count = 0;
  acu_status st;
  const int64_t nc = acu_resolve_null_count(ctx, a, &st);
  ACU_TRY(st);
  if (a->validity && nc > 0 && plan->count > 0) {
    const int grid = acu_grid(ctx, (plan->n_tiles + 7) / 8, 8);
    ACU_CUDA(ctx, cudaMemsetAsync(out->validity, 0, acu_bitmap_bytes(plan->count), ctx->stream));
    ACU_LAUNCH(ctx, k_filter_bits, grid, 256, 0, a->validity, a->validity_offset, plan->mask, plan->tile_local,
               plan->tile_count, plan->chunk_offset, plan->n_tiles, plan->len, reinterpret_cast<uint32_t *>(out->validity));
    ACU_TRY(acu_res_reset(ctx));
    ACU_TRY(acu_bitmap_and_launch(ctx, out->validity, 0, nullptr, 0, plan->count, nullptr, true));
    ACU_TRY(acu_res_fetch(ctx));
    const int64_t null_count = plan->count - (int64_t)ctx->h_res[RES_COUNT];
    if (null_count > 0) { out->has_validity = 1; out->null_count = null_count; }
  } else {
    ACU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return ACU_OK;
}

// Internal accessors for bytes.cu (filter_bytes materialises the selected row indices).
const uint64_t *acu_plan_mask(const acu_filter_plan *p) { return p->mask; }
const uint32_t *acu_plan_tile_local(const acu_filter_plan *p) { return p->tile_local; }
const uint32_t *acu_plan_tile_count(const acu_filter_plan *p) { return p->tile_count; }
const uint64_t *acu_plan_chunk_offset(const acu_filter_plan *p) { return p->chunk_offset; }
int64_t acu_plan_n_tiles(const acu_filter_plan *p) { return p->n_tiles; }
