/*
 * arrow_cuda.h — C ABI of the B200-native arrow::compute hot path.
 *
 * This is the drop-in boundary: one extern "C" entry point per
 * reference kernel, taking raw DEVICE pointers + explicit lengths / bit offsets, so
 * that a thin Rust `arrow-cuda` crate (rust/arrow-cuda, source only — no rustc in
 * this image) or the C++ host mirror (arrow-rs_b200/host/arrow_cuda.hpp) can wrap
 * them with the reference's own signatures over ArrayRef / RecordBatch.
 *
 * Layout contract (mirrors arrow-buffer; reference file:line in brackets):
 *   - values buffer: contiguous little-endian natives, pointer already advanced to
 *     logical element 0                      [arrow-buffer/src/buffer/scalar.rs:29-46]
 *   - validity / boolean bitmap: bytes, bit i of the logical array at byte
 *     (off+i)>>3, bit (off+i)&7, LSB first, 1 = valid / true
 *                                            [arrow-buffer/src/util/bit_util.rs:52-66,
 *                                             arrow-buffer/src/buffer/boolean.rs:97-104]
 *   - null_count is cached beside the bitmap  [arrow-buffer/src/buffer/null.rs:34-37]
 *
 * Every OUTPUT bitmap is written with bit offset 0 and must have a capacity of
 * acu_bitmap_bytes(len) = 8*ceil(len/64) bytes (kernels store whole 64-bit words);
 * bits at positions >= len are unspecified, exactly as in the reference
 * (arrow-ord/src/cmp.rs:598-608).
 *
 * Calls are synchronous with respect to the host: every entry point that returns a
 * host-visible scalar (count, null_count, error index) synchronises the ctx stream
 * before returning (stream-ordered sections, acu_async_begin below, defer that to ONE
 * synchronisation for a chain of calls). One acu_ctx = one device + one stream + scratch; a ctx must not
 * be used from two host threads at once, distinct ctxs are independent (the
 * reference kernels are pure, re-entrant functions: arrow-array/src/array/mod.rs:99).
 *
 * There is NO CPU fallback behind this ABI: if no CUDA device is present
 * acu_ctx_create fails with ACU_ERR_CUDA.
 */
#ifndef ARROW_CUDA_H
#define ARROW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACU_ABI_VERSION 1

/* ------------------------------------------------------------------------- */
/* Status codes — one per ArrowError variant the hot path can produce        */
/* (arrow-schema/src/error.rs:26-67), plus the reference's panics, plus      */
/* device-side failures.                                                     */
/* ------------------------------------------------------------------------- */
typedef int32_t acu_status;
enum {
  ACU_OK = 0,
  ACU_ERR_INVALID_ARGUMENT = 1,    /* ArrowError::InvalidArgumentError          */
  ACU_ERR_COMPUTE = 2,             /* ArrowError::ComputeError                  */
  ACU_ERR_ARITHMETIC_OVERFLOW = 3, /* ArrowError::ArithmeticOverflow            */
  ACU_ERR_DIVIDE_BY_ZERO = 4,      /* ArrowError::DivideByZero                  */
  ACU_ERR_OFFSET_OVERFLOW = 5,     /* ArrowError::OffsetOverflowError(usize)    */
  ACU_ERR_CAST = 6,                /* ArrowError::CastError                     */
  ACU_ERR_NOT_YET_IMPLEMENTED = 7, /* ArrowError::NotYetImplemented             */
  ACU_ERR_PANIC_OUT_OF_BOUNDS = 8, /* the reference panics (take.rs:447,454)    */
  ACU_ERR_IPC = 9,                 /* ArrowError::IpcError                      */
  ACU_ERR_PARSE = 10,              /* ArrowError::ParseError                    */
  ACU_ERR_CUDA = 100,              /* CUDA runtime error (detail.cuda_error)    */
  ACU_ERR_NCCL = 101,              /* NCCL error                                */
  ACU_ERR_OUT_OF_MEMORY = 102
};

/* Filled on every non-OK return; lets the host shim rebuild the reference's exact
 * message, e.g. "Overflow happened on: {lhs} + {rhs}" (arrow-array/src/arithmetic.rs:163-170)
 * or "Array index out of bounds, cannot get item at index {index} from {len} entries"
 * (arrow-select/src/take.rs:186-188). `message` already holds that text. */
typedef struct acu_error_detail {
  acu_status status;
  int32_t cuda_error;   /* cudaError_t / ncclResult_t when status >= 100 */
  int64_t index;        /* lowest offending logical row, -1 if n/a        */
  uint64_t lhs_bits;    /* operand bit patterns at `index`                */
  uint64_t rhs_bits;
  uint64_t len;         /* array length / capacity relevant to the error  */
  char message[256];
} acu_error_detail;

typedef struct acu_ctx acu_ctx;

/* Native types (arrow-array/src/types.rs:67-80, ArrowPrimitiveType::Native). */
typedef enum acu_dtype {
  ACU_I8 = 0, ACU_I16 = 1, ACU_I32 = 2, ACU_I64 = 3,
  ACU_U8 = 4, ACU_U16 = 5, ACU_U32 = 6, ACU_U64 = 7,
  ACU_F32 = 8, ACU_F64 = 9
} acu_dtype;

/* arrow-arith/src/numeric.rs:181-190 `enum Op` — same order. */
typedef enum acu_arith_op {
  ACU_ADD_WRAPPING = 0, ACU_ADD = 1,
  ACU_SUB_WRAPPING = 2, ACU_SUB = 3,
  ACU_MUL_WRAPPING = 4, ACU_MUL = 5,
  ACU_DIV = 6, ACU_REM = 7
} acu_arith_op;

/* arrow-ord/src/cmp.rs:40-60 `enum Op`. */
typedef enum acu_cmp_op {
  ACU_EQ = 0, ACU_NEQ = 1, ACU_LT = 2, ACU_LT_EQ = 3, ACU_GT = 4, ACU_GT_EQ = 5,
  ACU_DISTINCT = 6, ACU_NOT_DISTINCT = 7
} acu_cmp_op;

/* arrow-arith/src/aggregate.rs:943,1012,1027. */
typedef enum acu_agg_op { ACU_SUM = 0, ACU_MIN = 1, ACU_MAX = 2 } acu_agg_op;

/* A borrowed, immutable view of a primitive / boolean array in HBM
 * (PrimitiveArray{values,nulls} arrow-array/src/array/primitive_array.rs:596-601;
 *  BooleanArray{values,nulls} arrow-array/src/array/boolean_array.rs:68-71). */
typedef struct acu_array {
  const void *values;       /* device ptr to logical element 0 (for boolean arrays: bitmap bytes) */
  int64_t values_offset;    /* boolean arrays only: bit offset of logical row 0 in `values`      */
  const uint8_t *validity;  /* device ptr to validity bytes, NULL = no NullBuffer                */
  int64_t validity_offset;  /* bit offset of logical row 0 in `validity`                         */
  int64_t len;              /* logical length                                                    */
  int64_t null_count;       /* cached null count; -1 = unknown (counted on device)               */
  int32_t is_scalar;        /* Datum::get().1 (arrow-array/src/scalar.rs:78-152): len must be 1  */
  int32_t reserved;
} acu_array;

/* Caller-owned output. `values` capacity: len*width bytes (boolean results:
 * acu_bitmap_bytes(len)); `validity` capacity: acu_bitmap_bytes(len).
 * On return: len, null_count, has_validity (0 => the reference returns nulls = None
 * and the validity buffer content is unspecified). */
typedef struct acu_array_out {
  void *values;
  uint8_t *validity;
  int64_t len;
  int64_t null_count;
  int32_t has_validity;
  int32_t reserved;
} acu_array_out;

static inline size_t acu_bitmap_bytes(int64_t len) { return (size_t)((len + 63) / 64) * 8; }

/* ------------------------------------------------------------------------- */
/* Context, memory, timing                                                   */
/* ------------------------------------------------------------------------- */
int32_t acu_abi_version(void);
/* sizeof() of the ABI structs as this library was compiled, for binding self-checks:
 * 0 acu_array, 1 acu_array_out, 2 acu_error_detail, 3 acu_column, 4 acu_column_out; -1 otherwise. */
int32_t acu_abi_sizeof(int32_t which);
acu_status acu_ctx_create(int32_t device, acu_ctx **out);
void acu_ctx_destroy(acu_ctx *ctx);
acu_status acu_ctx_sync(acu_ctx *ctx);
/* Detail of the last failing call on this ctx (never NULL once ctx exists). */
const acu_error_detail *acu_last_error(const acu_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py: gpu_launches). */
int64_t acu_launch_count(const acu_ctx *ctx);
int32_t acu_device_sm_count(const acu_ctx *ctx);

/* DeviceBuffer allocation: 256-B aligned, stream-ordered (cudaMallocAsync pool);
 * mirrors arrow-buffer's 128-B aligned host allocation (src/alloc/alignment.rs:38). */
acu_status acu_malloc(acu_ctx *ctx, size_t bytes, void **out_dptr);
acu_status acu_free(acu_ctx *ctx, void *dptr);
acu_status acu_memset(acu_ctx *ctx, void *dptr, int32_t byte, size_t bytes);
acu_status acu_memcpy_h2d(acu_ctx *ctx, void *dst_dptr, const void *src_host, size_t bytes);
acu_status acu_memcpy_d2h(acu_ctx *ctx, void *dst_host, const void *src_dptr, size_t bytes);
acu_status acu_memcpy_d2d(acu_ctx *ctx, void *dst_dptr, const void *src_dptr, size_t bytes);
/* Asynchronous variants (ordered on the ctx stream; host memory must be pinned). */
acu_status acu_memcpy_h2d_async(acu_ctx *ctx, void *dst_dptr, const void *src_host, size_t bytes);
acu_status acu_memcpy_d2h_async(acu_ctx *ctx, void *dst_host, const void *src_dptr, size_t bytes);
acu_status acu_host_alloc(acu_ctx *ctx, size_t bytes, void **out_host);   /* pinned */
acu_status acu_host_free(acu_ctx *ctx, void *host);
/* Bytes currently allocated through acu_malloc on this ctx (observability; the
 * reference's MemoryPool tracking, arrow-buffer/src/pool.rs:73-85). */
int64_t acu_bytes_allocated(const acu_ctx *ctx);

/* ---- Stream-ordered sections -------------------------------------------------------------------------------------
 * The reference functions are synchronous (arrow-array/src/array/mod.rs:99) and so is every entry point by default. A
 * caller that chains several kernels (filter -> take -> add -> sum) can instead open a SECTION: between acu_async_begin
 * and acu_results_fetch the entry points listed below only ENQUEUE their kernels on the ctx stream and return at once;
 * each call owns one of the ctx's 64 result blocks (count, null count, lowest failing row stay in HBM).
 * acu_results_fetch copies all blocks to the host in ONE transfer with ONE synchronisation, then finalises the queued
 * calls in call order: it fills the acu_array_out / scalar outputs the calls were given (len, has_validity, null_count,
 * aggregate bits), fills acu_filter_plan count / strategy, and returns the FIRST error in call order with its exact
 * reference text (the outputs of the calls after a failed one are unspecified, as after any failed call).
 *   - stream-ordered inside a section: acu_filter_plan_create, acu_filter_plan_create_cmp, acu_filter_primitive,
 *     acu_filter_boolean, acu_take_primitive / acu_take_boolean (check_bounds = 0), acu_arith, acu_cmp, acu_aggregate,
 *     acu_aggregate_allreduce. Any other entry point fails with ACU_ERR_INVALID_ARGUMENT (it would synchronise).
 *   - every output descriptor, scalar output pointer and plan passed to a queued call must stay alive until the fetch;
 *     input arrays must carry their cached null_count (-1 would need a device count = a synchronisation), except for
 *     acu_aggregate*, which then counts the valid rows on the device (an array produced earlier in the same section);
 *   - a plan created inside the section can be used by filters queued after it: their outputs must then be sized for
 *     plan LEN rows (the count is still on the device); out->len is set by the fetch;
 *   - at most 64 calls per section. */
acu_status acu_async_begin(acu_ctx *ctx);
acu_status acu_results_fetch(acu_ctx *ctx);
int32_t acu_async_active(const acu_ctx *ctx);

/* CUDA-event timers on the ctx stream (events see exactly the stream kernels run on).
 * ACU_TIMER_SLOTS independent slots so that a step timer can bracket per-op timers. */
#define ACU_TIMER_SLOTS 8
acu_status acu_timer_start(acu_ctx *ctx);                /* slot 0 */
acu_status acu_timer_stop(acu_ctx *ctx, float *out_ms);  /* slot 0; records + synchronises */
acu_status acu_timer_start_slot(acu_ctx *ctx, int32_t slot);
acu_status acu_timer_stop_slot(acu_ctx *ctx, int32_t slot, float *out_ms);

/* Per-kernel device time, always on: every launch of a hot kernel is bracketed by a pair of
 * CUDA events on the ctx stream and its elapsed time is accumulated per kernel class when
 * the call synchronises. bench.py derives roofline.achieved from these (kernel-only) times. */
typedef enum acu_kernel_class {
  ACU_K_ARITH = 0, ACU_K_CMP = 1, ACU_K_CAST = 2, ACU_K_FILTER = 3, ACU_K_FILTER_PLAN = 4,
  ACU_K_TAKE = 5, ACU_K_REDUCE = 6, ACU_K_BYTES = 7, ACU_K_CLASSES = 8
} acu_kernel_class;
acu_status acu_kernel_stats(acu_ctx *ctx, int32_t kernel_class, double *out_total_ms, int64_t *out_launches);
acu_status acu_kernel_stats_reset(acu_ctx *ctx);


/* ------------------------------------------------------------------------- */
/* Bitmaps                                                                   */
/* ------------------------------------------------------------------------- */
/* popcount of bits [offset, offset+len)  — BooleanBuffer::count_set_bits
 * (arrow-buffer/src/buffer/boolean.rs) / BooleanArray::true_count when `validity`
 * is non-NULL (arrow-array/src/array/boolean_array.rs:175-187). */
acu_status acu_bitmap_count(acu_ctx *ctx, const uint8_t *bits, int64_t offset,
                            const uint8_t *validity, int64_t validity_offset, int64_t len,
                            int64_t *out_count);

/* ------------------------------------------------------------------------- */
/* filter — arrow-select/src/filter.rs                                       */
/* ------------------------------------------------------------------------- */
/* FilterBuilder::new + optimize + build (filter.rs:254-324): folds predicate nulls
 * into the mask (prep_null_mask_filter :167-171), counts selected rows, and keeps a
 * device-resident plan (normalised mask words + per-tile output offsets) that can be
 * applied to any number of columns (FilterPredicate, filter.rs:442-533). */
typedef struct acu_filter_plan acu_filter_plan;
typedef enum acu_filter_strategy {  /* IterationStrategy, filter.rs:328-365 */
  ACU_FILTER_NONE = 0, ACU_FILTER_ALL = 1, ACU_FILTER_INDEX = 2, ACU_FILTER_SLICES = 3
} acu_filter_strategy;

acu_status acu_filter_plan_create(acu_ctx *ctx, const acu_array *predicate /* boolean array */,
                                  acu_filter_plan **out_plan);
void acu_filter_plan_destroy(acu_ctx *ctx, acu_filter_plan *plan);
/* FilterBuilder::optimize (filter.rs:285-298): materialise IterationStrategy::Indices — the
 * selected row ids in ascending order, as index_dtype ACU_U32 or ACU_U64 (count entries). */
acu_status acu_filter_plan_indices(acu_ctx *ctx, const acu_filter_plan *plan, acu_dtype index_dtype,
                                   void *out_indices);
/* FilterBuilder::optimize for dense predicates: IterationStrategy::Slices(Vec<(usize, usize)>) = SlicesIterator
 * (filter.rs:44-77, :285-298): the runs of selected rows as [start, end) pairs in ascending order, out_pairs[2k] = start,
 * out_pairs[2k + 1] = end of run k. *out_slices = number of runs; out_pairs == NULL sizes only. Null predicate slots select
 * nothing (the plan's mask is prep_null_mask_filter'ed). */
acu_status acu_filter_plan_slices(acu_ctx *ctx, const acu_filter_plan *plan, uint64_t *out_pairs, int64_t capacity,
                                  int64_t *out_slices);
int64_t acu_filter_plan_count(const acu_filter_plan *plan);      /* FilterPredicate::count */
int64_t acu_filter_plan_len(const acu_filter_plan *plan);        /* predicate length       */
int32_t acu_filter_plan_strategy(const acu_filter_plan *plan);   /* acu_filter_strategy    */

/* filter_primitive / filter_native + filter_nulls (filter.rs:732-788, :512-533).
 * elem_bytes in {1,2,4,8,16,32}. Error if plan len > values.len (filter.rs:536-542).
 * out->values capacity: count*elem_bytes. out->has_validity = 0 when the source has no
 * nulls or the filtered result has none (filter.rs:518-526). */
acu_status acu_filter_primitive(acu_ctx *ctx, const acu_filter_plan *plan, int32_t elem_bytes,
                                const acu_array *values, acu_array_out *out);
/* filter_boolean / filter_bits (filter.rs:680-729): `values` is a boolean array. */
acu_status acu_filter_boolean(acu_ctx *ctx, const acu_filter_plan *plan, const acu_array *values,
                              acu_array_out *out);
/* filter_bytes for Utf8/Binary (offset_bytes = 4) and Large* (8) (filter.rs:893-928).
 * Two-phase: out_offsets (count+1 entries) is always written and *out_data_len returned;
 * value bytes are copied only if out_data != NULL (capacity out_data_capacity). */
acu_status acu_filter_bytes(acu_ctx *ctx, const acu_filter_plan *plan, int32_t offset_bytes,
                            const void *offsets, const uint8_t *data, const acu_array *nulls_of,
                            void *out_offsets, uint8_t *out_data, int64_t out_data_capacity,
                            int64_t *out_data_len, acu_array_out *out_nulls);

/* FilterBuilder::new(&cmp::op(a, b)?) without the BooleanArray in between (arrow-ord/src/cmp.rs:220-382
 * feeding arrow-select/src/filter.rs:254-273): the comparison writes the plan's normalised mask (result & result validity —
 * a null comparison result selects nothing, prep_null_mask_filter filter.rs:167-171) and the per-tile counts directly;
 * the 2 x N/8-byte result bitmaps are never written to or re-read from HBM. The plan is identical to
 * acu_filter_plan_create(acu_cmp(dtype, op, a, b)). Errors as acu_cmp. */
acu_status acu_filter_plan_create_cmp(acu_ctx *ctx, acu_dtype dtype, acu_cmp_op op, const acu_array *a,
                                      const acu_array *b, acu_filter_plan **out_plan);

/* ------------------------------------------------------------------------- */
/* nullif / zip — arrow-select/src/nullif.rs, zip.rs                         */
/* ------------------------------------------------------------------------- */
/* nullif(left, right) (nullif.rs:44-113): the result shares left's value buffers; only the validity changes:
 * out->validity = left.validity & !(right.values & right.validity) (bit offset 0), out->null_count, out->has_validity = 0
 * when no slot is null (ArrayDataBuilder::build drops an all-valid NullBuffer). `left` may be an array of any kind: only
 * its validity / len are read; out->values is not touched. Length mismatch => ACU_ERR_COMPUTE. */
acu_status acu_nullif(acu_ctx *ctx, const acu_array *left, const acu_array *right /* boolean */, acu_array_out *out);

/* zip(mask, truthy, falsy) (zip.rs:99-226) for fixed-width values of elem_bytes in {1,2,4,8,16,32}: out[i] = truthy[i] where
 * mask[i] is Some(true), else falsy[i]; either side may be a scalar (is_scalar, len 1); value bytes are copied blindly from
 * the chosen side (also under nulls); the result carries a validity buffer iff some input has nulls and the result has at
 * least one. Both sides scalar = ScalarZipper (zip.rs:248-440): with one null scalar every slot holds the other value and
 * the validity (always present) is the mask / its negation. ACU_ERR_INVALID_ARGUMENT "all arrays should have the same
 * length" / "scalar arrays must have 1 element". */
acu_status acu_zip(acu_ctx *ctx, int32_t elem_bytes, const acu_array *mask /* boolean */, const acu_array *truthy,
                   const acu_array *falsy, acu_array_out *out);

/* ------------------------------------------------------------------------- */
/* take — arrow-select/src/take.rs                                           */
/* ------------------------------------------------------------------------- */
/* take_primitive = take_native + take_nulls (take.rs:405-457). `indices.values`
 * has native type index_dtype (any integer type; ToIndices take.rs:1030-1084:
 * i8/i16 sign-extend to u32, i32/i64 reinterpret). check_bounds != 0 =>
 * TakeOptions{check_bounds:true} (take.rs:167-209): ACU_ERR_COMPUTE with the lowest
 * offending index. Otherwise an out-of-bounds VALID index returns
 * ACU_ERR_PANIC_OUT_OF_BOUNDS (the reference panics) and an out-of-bounds index in a
 * NULL slot yields T::default() = 0 (take.rs:442-448). */
acu_status acu_take_primitive(acu_ctx *ctx, int32_t elem_bytes, const acu_array *values,
                              const acu_array *indices, acu_dtype index_dtype,
                              int32_t check_bounds, acu_array_out *out);
/* take_boolean / take_bits (take.rs:460-496). */
acu_status acu_take_boolean(acu_ctx *ctx, const acu_array *values, const acu_array *indices,
                            acu_dtype index_dtype, int32_t check_bounds, acu_array_out *out);
/* take_bytes (take.rs:499-627); also Dictionary<K,Utf8> -> Utf8 cast =
 * unpack_dictionary (arrow-cast/src/cast/dictionary.rs:310-317) with values = the
 * dictionary and indices = the keys. i32 offset overflow => ACU_ERR_OFFSET_OVERFLOW
 * (take.rs:520-523). Two-phase like acu_filter_bytes. `nulls_of` carries the
 * validity/len/null_count of the byte array (its `values` member is ignored). */
acu_status acu_take_bytes(acu_ctx *ctx, int32_t offset_bytes, const void *offsets,
                          const uint8_t *data, const acu_array *nulls_of,
                          const acu_array *indices, acu_dtype index_dtype, int32_t check_bounds,
                          void *out_offsets, uint8_t *out_data, int64_t out_data_capacity,
                          int64_t *out_data_len, acu_array_out *out_nulls);

/* ------------------------------------------------------------------------- */
/* numeric — arrow-arith/src/numeric.rs, arity.rs                            */
/* ------------------------------------------------------------------------- */
/* add/add_wrapping/sub/sub_wrapping/mul/mul_wrapping/div/rem (numeric.rs:36-81).
 * a and b have native type `dtype`; either may be a scalar (is_scalar, len 1).
 * Floats: every op is the IEEE single operation (no FMA contraction), never errors.
 * Integers: wrapping ops via `binary` (op evaluated at every slot), checked ops via
 * `try_binary` (zero under nulls, first failing valid index reported)
 * (arity.rs:104-135, :254-299). Length mismatch => ACU_ERR_COMPUTE.
 * In place (binary_mut / try_binary_mut / unary_mut / try_unary_mut, arity.rs:137-252,301-363): out->values may BE a->values
 * (or b->values), and out->validity an input validity buffer whose bit offset is 0 — every element / bitmap word is read
 * before the same element / word is written. A failed checked op leaves the buffer partially overwritten (the reference's
 * try_binary_mut consumes its input as well). */
acu_status acu_arith(acu_ctx *ctx, acu_dtype dtype, acu_arith_op op, const acu_array *a,
                     const acu_array *b, acu_array_out *out);
/* neg (checked != 0) / neg_wrapping (numeric.rs:103-186). */
acu_status acu_neg(acu_ctx *ctx, acu_dtype dtype, int32_t checked, const acu_array *a,
                   acu_array_out *out);

/* ------------------------------------------------------------------------- */
/* cmp — arrow-ord/src/cmp.rs                                                */
/* ------------------------------------------------------------------------- */
/* eq/neq/lt/lt_eq/gt/gt_eq/distinct/not_distinct (cmp.rs:79-202). Result is a boolean
 * array: out->values = bit-packed results. Floats compare by IEEE-754 totalOrder
 * (arrow-array/src/arithmetic.rs:400-410). */
acu_status acu_cmp(acu_ctx *ctx, acu_dtype dtype, acu_cmp_op op, const acu_array *a,
                   const acu_array *b, acu_array_out *out);

/* The same eight comparisons on variable-width operands.
 * acu_bytes_array = GenericByteArray (Utf8 / Binary: offset_bytes 4, LargeUtf8 / LargeBinary: 8; ArrayOrd cmp.rs:783-801):
 * `offsets` points at the offset of logical row 0 (nulls.len + 1 entries), `nulls` carries len / validity / null_count /
 * is_scalar (its `values` member is ignored). acu_view_array = GenericByteViewArray (Utf8View / BinaryView; ArrayOrd
 * cmp.rs:803-898): `views` = 16 bytes per row (length u32, then 12 inline bytes, or 4-byte prefix + buffer index u32 + offset
 * u32: arrow-data/src/byte_view.rs), `buffers` = a HOST array of n_buffers DEVICE pointers to the data buffers. Bytes compare
 * like Rust's `&[u8]` (lexicographic on unsigned bytes, then length). Equality of a view array against a non-null constant of
 * at most 4 bytes takes the reference's short-constant path (eq_inline_scalar, cmp.rs:405-435: one masked 64-bit compare per
 * view). Null handling, result layout and errors exactly as acu_cmp. */
typedef struct acu_bytes_array {
  const void *offsets;
  const uint8_t *data;
  acu_array nulls;
} acu_bytes_array;
typedef struct acu_view_array {
  const void *views;
  const uint8_t *const *buffers;
  int32_t n_buffers;
  int32_t reserved;
  acu_array nulls;
} acu_view_array;
acu_status acu_cmp_bytes(acu_ctx *ctx, int32_t offset_bytes, acu_cmp_op op, const acu_bytes_array *l,
                         const acu_bytes_array *r, acu_array_out *out);
acu_status acu_cmp_byte_view(acu_ctx *ctx, acu_cmp_op op, const acu_view_array *l, const acu_view_array *r,
                             acu_array_out *out);

/* Utf8View / BinaryView buffer management for BatchCoalescer (InProgressByteViewArray, arrow-select/src/coalesce/
 * byte_view.rs). The reference decides per source array whether its data buffers are compacted ("gc": when they hold more
 * than twice the bytes its views use, :366-381) and how output buffers are sized (BufferSource, :526-559); that policy stays
 * on the host (host/arrow_cuda.hpp, acu/coalesce.py). The per-view work runs on the device:
 *   acu_view_bytes_used   = GenericByteViewArray::total_buffer_bytes_used (arrow-array/src/array/byte_view_array.rs:749-761):
 *                           sum of the lengths of the views longer than 12 bytes (null slots included, as in the reference).
 *   acu_view_fit          = the "copy as many views as fit the current buffer" loop (:259-271): *out_views = leading views
 *                           that fit `remaining_capacity` (EVERY view's length is compared with what is left, only the long
 *                           ones consume it — the reference's loop), *out_bytes = bytes of the long views among them.
 *   acu_view_copy_strings = append_views_and_copy_strings_inner (:298-354): out_views[i] = views[i], every view longer than
 *                           12 bytes rewritten to {buffer_index = new_buffer_index, offset = position in dst} with its bytes
 *                           copied to dst[dst_len ...] in view order (null slots too); *out_bytes = bytes appended. `buffers` =
 *                           HOST array of n_buffers DEVICE pointers (the source's data buffers).
 *   acu_view_rebase       = append_views_and_update_buffer_index (:176-216): buffer_index += delta for the long views.
 * views / out_views: 16 bytes per row, 16-byte aligned; out_views may alias views. */
acu_status acu_view_bytes_used(acu_ctx *ctx, const void *views, int64_t n, int64_t *out_total);
acu_status acu_view_fit(acu_ctx *ctx, const void *views, int64_t n, int64_t remaining_capacity, int64_t *out_views,
                        int64_t *out_bytes);
acu_status acu_view_copy_strings(acu_ctx *ctx, const void *views, int64_t n, const uint8_t *const *buffers,
                                 int32_t n_buffers, uint32_t new_buffer_index, uint8_t *dst, int64_t dst_len,
                                 int64_t dst_capacity, void *out_views, int64_t *out_bytes);
acu_status acu_view_rebase(acu_ctx *ctx, const void *views, int64_t n, uint32_t delta, void *out_views);

/* ------------------------------------------------------------------------- */
/* cast — arrow-cast/src/cast/mod.rs                                         */
/* ------------------------------------------------------------------------- */
/* cast_numeric_arrays (mod.rs:2550-2614). safe != 0 => numeric_cast (unrepresentable
 * value -> null, output always has a validity buffer: primitive_array.rs:1065-1103);
 * safe == 0 => try_numeric_cast (ACU_ERR_CAST "Can't cast value {v} to type {T}"). */
acu_status acu_cast_numeric(acu_ctx *ctx, acu_dtype from, acu_dtype to, int32_t safe,
                            const acu_array *a, acu_array_out *out);

/* ------------------------------------------------------------------------- */
/* boolean — arrow-arith/src/boolean.rs (predicate construction before filter) */
/* ------------------------------------------------------------------------- */
typedef enum acu_bool_op {
  ACU_BOOL_AND = 0,         /* and        boolean.rs:256  values a&b at every slot, nulls = union */
  ACU_BOOL_OR = 1,          /* or         boolean.rs:273 */
  ACU_BOOL_AND_NOT = 2,     /* and_not    boolean.rs:291  a & !b */
  ACU_BOOL_AND_KLEENE = 3,  /* and_kleene boolean.rs:60   false AND null = false */
  ACU_BOOL_OR_KLEENE = 4,   /* or_kleene  boolean.rs:156  true OR null = true */
  ACU_BOOL_NOT = 5,         /* not        boolean.rs:310  (b == NULL) */
  ACU_BOOL_IS_NULL = 6,     /* is_null    boolean.rs:327  any array kind: only a's validity/len are read; b == NULL */
  ACU_BOOL_IS_NOT_NULL = 7  /* is_not_null boolean.rs:347 */
} acu_bool_op;
/* a, b: BooleanArrays (values = bitmaps with bit offsets). out->values receives the result
 * bitmap (bit offset 0, whole u64 words, bits >= len zero). The result carries a validity
 * buffer exactly when the reference's does: and/or/and_not/kleene when either input has
 * one (even without nulls), not when `a` has one, is_null/is_not_null never. Length
 * mismatch => ACU_ERR_COMPUTE "Cannot perform bitwise operation on arrays of different
 * length". */
acu_status acu_boolean(acu_ctx *ctx, acu_bool_op op, const acu_array *a, const acu_array *b,
                       acu_array_out *out);

/* ------------------------------------------------------------------------- */
/* aggregate — arrow-arith/src/aggregate.rs                                  */
/* ------------------------------------------------------------------------- */
/* sum/min/max (aggregate.rs:943,1012,1027): *out_bits = the native result's bit
 * pattern (zero-extended), *out_valid_count = number of non-null rows; the reference
 * returns None iff out_valid_count == 0 (aggregate.rs:320-323). */
acu_status acu_aggregate(acu_ctx *ctx, acu_dtype dtype, acu_agg_op op, const acu_array *a,
                         uint64_t *out_bits, int64_t *out_valid_count);

/* sum_checked (aggregate.rs:897-937): the in-order checked fold. Integers:
 * ACU_ERR_ARITHMETIC_OVERFLOW "Overflow happened on: {acc} + {value}" at the first valid
 * row whose running sum leaves the native range (index = that row) — also when the final
 * total would fit. Floats never fail (add_checked is the plain add): same as ACU_SUM. */
acu_status acu_sum_checked(acu_ctx *ctx, acu_dtype dtype, const acu_array *a, uint64_t *out_bits,
                           int64_t *out_valid_count);

/* ------------------------------------------------------------------------- */
/* RecordBatch level — filter_record_batch / take_record_batch / per-column   */
/* aggregates with ONE stream synchronisation per call                        */
/* ------------------------------------------------------------------------- */
/* A column of a RecordBatch (arrow-array/src/record_batch.rs:224). PRIMITIVE: `array`
 * as for acu_filter_primitive with element width `width`; BOOLEAN: `array.values` is a
 * bitmap; BYTES (Utf8/Binary/LargeUtf8/LargeBinary): `array.values` = the offsets
 * buffer (i32 if width == 4, i64 if width == 8, array.len + 1 entries), `data` = the
 * value bytes, `array.validity/len/null_count` the nulls. */
typedef enum acu_column_kind { ACU_COL_PRIMITIVE = 0, ACU_COL_BOOLEAN = 1, ACU_COL_BYTES = 2 } acu_column_kind;

typedef struct acu_column {
  int32_t kind;          /* acu_column_kind */
  int32_t width;         /* PRIMITIVE: element bytes (1,2,4,8,16,32); BYTES: offset bytes (4|8) */
  acu_array array;
  const uint8_t *data;   /* BYTES only */
} acu_column;

/* Caller-owned output of one column. `array.values` receives the values (PRIMITIVE /
 * BOOLEAN) or the new offsets (BYTES, rows + 1 entries); `data` (capacity
 * `data_capacity`) the value bytes; `data_len` is set to the bytes required/written. */
typedef struct acu_column_out {
  acu_array_out array;
  uint8_t *data;
  int64_t data_capacity;
  int64_t data_len;
} acu_column_out;

/* filter_record_batch (filter.rs:225-244) / FilterPredicate::filter_record_batch
 * (filter.rs:459-478): every column filtered with the same plan (the predicate is
 * scanned once), all kernels queued back to back, one synchronisation. Columns keep
 * their order; on error the status/detail of the first failing column is returned (the
 * reference propagates that column's ArrowError). At most ACU_MAX_BATCH_COLUMNS columns. */
#define ACU_MAX_BATCH_COLUMNS 64
acu_status acu_filter_record_batch(acu_ctx *ctx, const acu_filter_plan *plan, int32_t n_columns,
                                   const acu_column *columns, acu_column_out *outs);

/* take_record_batch (take.rs:1123-1133) / take_arrays (take.rs:155-164): every column
 * gathered with the same indices; check_bounds as in acu_take_primitive. Null counts of
 * the columns and of the indices should be cached (>= 0); unknown ones are counted
 * first (one extra synchronisation each). */
acu_status acu_take_record_batch(acu_ctx *ctx, int32_t n_columns, const acu_column *columns,
                                 const acu_array *indices, acu_dtype index_dtype, int32_t check_bounds,
                                 acu_column_out *outs);

/* sum/min/max of n columns (acu_aggregate semantics per column) with one
 * synchronisation: dtypes[i]/ops[i] describe arrays[i]. */
acu_status acu_aggregate_columns(acu_ctx *ctx, int32_t n_columns, const acu_dtype *dtypes,
                                 const acu_agg_op *ops, const acu_array *arrays, uint64_t *out_bits,
                                 int64_t *out_valid_counts);

/* ------------------------------------------------------------------------- */
/* appending row ranges — the pieces of BatchCoalescer / concat               */
/* (arrow-select/src/coalesce.rs:258-533 InProgressArray::copy_rows)          */
/* ------------------------------------------------------------------------- */
/* Values of a fixed-width column append with acu_memcpy_d2d. Bits (validity, boolean
 * values): dst bits [dst_offset, dst_offset+len) = src bits [src_offset, ..+len), every
 * other bit of dst preserved (dst 8-byte aligned, capacity a whole number of u64 words).
 * *out_set_bits (optional; forces a synchronisation) = number of set bits copied. */
acu_status acu_bitmap_copy(acu_ctx *ctx, const uint8_t *src, int64_t src_offset, uint8_t *dst,
                           int64_t dst_offset, int64_t len, int64_t *out_set_bits);
/* dst bits [dst_offset, dst_offset+len) = value (0 | 1), other bits preserved. */
acu_status acu_bitmap_fill(acu_ctx *ctx, uint8_t *dst, int64_t dst_offset, int64_t len, int32_t value);
/* Utf8/Binary offsets of `count` rows starting at source row `first`, rebased so that the
 * first one equals `base` (the destination's byte total so far):
 * dst[dst_first + j] = base + src[first + j] - src[first], j = 0..count. Returns the source
 * byte range [*out_src_begin, *out_src_end) to append with acu_memcpy_d2d.
 * ACU_ERR_OFFSET_OVERFLOW when the running total leaves the offset type. */
acu_status acu_offsets_append(acu_ctx *ctx, int32_t offset_bytes, const void *src_offsets, int64_t first,
                              int64_t count, int64_t base, void *dst_offsets, int64_t dst_first,
                              int64_t *out_src_begin, int64_t *out_src_end);

/* concat (arrow-select/src/concat.rs:495-577): n columns of the same kind / width appended in order into the caller-owned
 * `out` (capacities: total rows * width; boolean / validity bitmaps acu_bitmap_bytes(total rows); BYTES: total rows + 1
 * offsets and out->data_capacity value bytes). Values and the bytes under null slots are copied as they are; the result
 * carries a validity buffer iff some input has nulls (NullBufferBuilder semantics) — a single input keeps its NullBuffer
 * presence (the reference returns array.slice(0, len)). Errors: no input => ACU_ERR_COMPUTE "concat requires input of at
 * least one array"; mixed kinds / widths => ACU_ERR_INVALID_ARGUMENT; i32 offsets overflowing => ACU_ERR_OFFSET_OVERFLOW
 * (generic_bytes_builder.rs:185-189). Inputs without a cached null_count cost one count each. */
acu_status acu_concat(acu_ctx *ctx, int32_t n_arrays, const acu_column *arrays, acu_column_out *out);
/* concat_batches (concat.rs:607-640): columns[b * n_columns + c] = column c of batch b; outs[c] = concat of field c over the
 * batches; *out_rows = rows of the result. No batch => every field is empty. */
acu_status acu_concat_batches(acu_ctx *ctx, int32_t n_batches, int32_t n_columns, const acu_column *columns,
                              acu_column_out *outs, int64_t *out_rows);

/* ------------------------------------------------------------------------- */
/* Arrow C Data Interface / C Device Data Interface                           */
/* ------------------------------------------------------------------------- */
/* The structs of the Arrow specification (the reference's FFI_ArrowArray / FFI_ArrowSchema,
 * arrow-data/src/ffi.rs:37-69, arrow-schema/src/ffi.rs). Guarded like the specification's
 * own header so that they can coexist with <arrow/c/abi.h>. */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
  const char *format;
  const char *name;
  const char *metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema **children;
  struct ArrowSchema *dictionary;
  void (*release)(struct ArrowSchema *);
  void *private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void **buffers;
  struct ArrowArray **children;
  struct ArrowArray *dictionary;
  void (*release)(struct ArrowArray *);
  void *private_data;
};
#endif
#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE
typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
#define ARROW_DEVICE_CUDA_HOST 3
struct ArrowDeviceArray {
  struct ArrowArray array;
  int64_t device_id;
  ArrowDeviceType device_type;
  void *sync_event;
  int64_t reserved[3];
};
#endif
/* Export a column (no copy): buffers[0] = validity, [1] = values | offsets, [2] = bytes, one
 * logical offset (= the validity / boolean bit offset; `values` pointers are stepped back by
 * it, so the column must be a slice of a buffer whose row 0 is addressable — true of every
 * array this library or Arrow produces). `dtype` names the primitive type for the schema
 * format. The consumer's call of out_array->array.release invokes release_owner(owner) exactly
 * once. ctx may be NULL only for ARROW_DEVICE_CPU columns. For device columns the ctx stream is
 * synchronised first and sync_event is NULL. On any error out_array->array.release is NULL (nothing to release). */
acu_status acu_export_column(acu_ctx *ctx, const acu_column *col, acu_dtype dtype, int32_t device_type,
                             void (*release_owner)(void *), void *owner,
                             struct ArrowDeviceArray *out_array, struct ArrowSchema *out_schema);
/* View an imported (device or host) array as an acu_column: pointers into the producer's
 * buffers, valid until the caller invokes in->array.release. Flat primitive / boolean /
 * (large) utf8 / binary formats; anything else => ACU_ERR_NOT_YET_IMPLEMENTED.
 * Device rules of the C Device Data Interface: for ARROW_DEVICE_CUDA / CUDA_HOST arrays `ctx` is required, a CUDA array's
 * device_id must be the ctx's device (ACU_ERR_INVALID_ARGUMENT otherwise), and a non-NULL sync_event (a cudaEvent_t*) is
 * waited on by the ctx stream before any later kernel of this ctx can touch the buffers. ARROW_DEVICE_CPU arrays yield HOST
 * pointers (in->device_type tells the caller; ctx may be NULL); other device types => ACU_ERR_NOT_YET_IMPLEMENTED. */
acu_status acu_import_column(acu_ctx *ctx, const struct ArrowDeviceArray *in, const struct ArrowSchema *schema,
                             acu_column *out, acu_dtype *out_dtype);

/* ------------------------------------------------------------------------- */
/* Arrow IPC stream -> HBM (arrow-ipc/src/reader.rs StreamReader)            */
/* ------------------------------------------------------------------------- */
/* StreamReader::try_new (reader.rs:1587-1640) over an in-memory IPC stream (`stream` must stay valid until close): reads the
 * schema message. Flat primitive / boolean / Utf8 / Binary / LargeUtf8 / LargeBinary fields, uncompressed little-endian
 * bodies; anything else => ACU_ERR_NOT_YET_IMPLEMENTED naming the field. Errors keep the reference's texts ("Expected schema
 * message, found empty stream.", "Expected a schema as the first message in the stream, got: RecordBatch", ...). */
typedef struct acu_ipc_stream acu_ipc_stream;
acu_status acu_ipc_stream_open(acu_ctx *ctx, const uint8_t *stream, int64_t stream_len, acu_ipc_stream **out,
                               int32_t *out_n_fields);
/* Field i of the schema: kind (acu_column_kind), width (element / offset bytes), dtype (acu_dtype, -1 for boolean and byte
 * fields), nullable, name (owned by the stream). */
acu_status acu_ipc_stream_field(const acu_ipc_stream *s, int32_t i, int32_t *kind, int32_t *width, int32_t *dtype,
                                int32_t *nullable, const char **name);
/* StreamReader::next (maybe_next, reader.rs:1646-1671): decodes the next RecordBatch message. Its body goes to HBM with ONE
 * host->device copy into a buffer owned by the stream, and out_columns[0..n_fields) are VIEWS into that buffer (IPC body
 * buffers are 8-byte aligned Arrow buffers: nothing is re-laid out); they stay valid until the next call or close. The
 * validity pointer is NULL when the field node's null_count is 0 (reader.rs:271). *out_rows = -1 at the end of the stream. */
acu_status acu_ipc_stream_next(acu_ctx *ctx, acu_ipc_stream *s, acu_column *out_columns, int64_t *out_rows);
void acu_ipc_stream_close(acu_ctx *ctx, acu_ipc_stream *s);

/* ------------------------------------------------------------------------- */
/* multi-GPU: row-range shards, NCCL only for the final scalar reduce        */
/* ------------------------------------------------------------------------- */
#define ACU_NCCL_UNIQUE_ID_BYTES 128
acu_status acu_comm_get_unique_id(uint8_t out_id[ACU_NCCL_UNIQUE_ID_BYTES]);
acu_status acu_comm_init(acu_ctx *ctx, const uint8_t id[ACU_NCCL_UNIQUE_ID_BYTES], int32_t rank,
                         int32_t world_size);
acu_status acu_comm_destroy(acu_ctx *ctx);
/* All-reduce `n` per-shard partial aggregates of one (dtype, op) in one NCCL call.
 * partial_bits[i] / valid_counts[i] are what acu_aggregate returned on this rank;
 * on return they hold the global result. Float min/max are reduced on their
 * totalOrder integer keys (NCCL's float min/max are IEEE, not totalOrder). */
acu_status acu_comm_allreduce_aggregates(acu_ctx *ctx, acu_dtype dtype, acu_agg_op op,
                                         uint64_t *partial_bits, int64_t *valid_counts,
                                         int32_t n);
/* acu_aggregate over this rank's shard combined over all ranks in ONE call with ONE synchronisation: the partial stays in
 * HBM, a one-thread kernel re-encodes it (identity for a shard without valid rows, totalOrder key for float / signed
 * min / max), NCCL reduces {value, valid_count} in place on the ctx stream, and only the final pair crosses to the host
 * (no host bounce between the reduction kernel and the collective). Without a communicator it is acu_aggregate. */
acu_status acu_aggregate_allreduce(acu_ctx *ctx, acu_dtype dtype, acu_agg_op op, const acu_array *a,
                                   uint64_t *out_bits, int64_t *out_valid_count);
/* Sum int64 scalars across ranks (row counts, null counts). */
acu_status acu_comm_allreduce_i64_sum(acu_ctx *ctx, int64_t *values, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* ARROW_CUDA_H */
