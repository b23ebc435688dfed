//! `arrow-cuda`: drop-in for the `arrow::compute` hot path, backed by hand-written sm_100a kernels
//! (libarrow_cuda.so). SOURCE ONLY in this repository — the compiled, tested equivalents of this
//! layer are `arrow-rs_b200/host/arrow_cuda.hpp` (C++) and `arrow-rs_b200/acu` (Python).
//!
//! The functions keep the reference signatures:
//!   `arrow_select::filter::filter`            (arrow-select/src/filter.rs:201)
//!   `arrow_select::take::take`                (arrow-select/src/take.rs:89)
//!   `arrow_arith::numeric::{add, ...}`        (arrow-arith/src/numeric.rs:36-81)
//!   `arrow_ord::cmp::{eq, ...}`               (arrow-ord/src/cmp.rs:79-202)
//!   `arrow_cast::cast`                        (arrow-cast/src/cast/mod.rs:347)
//!   `arrow_arith::aggregate::{sum, min, max}` (arrow-arith/src/aggregate.rs:943,1012,1027)
pub mod ffi;

use arrow_array::{Array, ArrayRef, BooleanArray, Datum, PrimitiveArray, ArrowPrimitiveType};
use arrow_buffer::{BooleanBuffer, Buffer, NullBuffer, ScalarBuffer};
use arrow_schema::{ArrowError, DataType};
use std::ffi::CStr;
use std::os::raw::c_void;
use std::sync::Arc;

/// One device + one stream (`acu_ctx`). The reference kernels are pure functions; the context is
/// the implicit "where does this run". Not `Sync`: use one per thread.
pub struct Context { raw: *mut ffi::acu_ctx }

impl Context {
    pub fn new(device: i32) -> Result<Self, ArrowError> {
        let mut raw = std::ptr::null_mut();
        match unsafe { ffi::acu_ctx_create(device, &mut raw) } {
            ffi::ACU_OK => Ok(Self { raw }),
            st => Err(ArrowError::ExternalError(format!("acu_ctx_create failed ({st}): no CUDA device, no CPU fallback").into())),
        }
    }
    /// Map a non-OK status to the ArrowError variant the reference would have returned; the
    /// message text is the reference's own (pinned by tests/golden/vectors.json).
    fn error(&self, st: ffi::acu_status) -> ArrowError {
        let d = unsafe { &*ffi::acu_last_error(self.raw) };
        let full = unsafe { CStr::from_ptr(d.message.as_ptr()) }.to_string_lossy().into_owned();
        let strip = |p: &str| full.strip_prefix(p).unwrap_or(&full).to_string();
        match st {
            ffi::ACU_ERR_INVALID_ARGUMENT => ArrowError::InvalidArgumentError(strip("Invalid argument error: ")),
            ffi::ACU_ERR_COMPUTE => ArrowError::ComputeError(strip("Compute error: ")),
            ffi::ACU_ERR_ARITHMETIC_OVERFLOW => ArrowError::ArithmeticOverflow(strip("Arithmetic overflow: ")),
            ffi::ACU_ERR_DIVIDE_BY_ZERO => ArrowError::DivideByZero,
            ffi::ACU_ERR_OFFSET_OVERFLOW => ArrowError::OffsetOverflowError(d.len as usize),
            ffi::ACU_ERR_CAST => ArrowError::CastError(strip("Cast error: ")),
            ffi::ACU_ERR_PANIC_OUT_OF_BOUNDS => panic!("{full}"), // take.rs:447 — the reference panics
            _ => ArrowError::ExternalError(full.into()),
        }
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::acu_ctx_destroy(self.raw) } } }

/// DeviceBuffer: mirrors `arrow_buffer::Buffer { data: Arc<Bytes>, ptr, length }`
/// (arrow-buffer/src/buffer/immutable.rs:83-96) with the bytes in HBM.
pub struct DeviceBuffer { ctx: *mut ffi::acu_ctx, ptr: *mut c_void, len: usize }
impl DeviceBuffer {
    pub fn from_host(ctx: &Context, bytes: &[u8]) -> Result<Arc<Self>, ArrowError> {
        let mut ptr = std::ptr::null_mut();
        let st = unsafe { ffi::acu_malloc(ctx.raw, bytes.len() + 16, &mut ptr) };
        if st != ffi::ACU_OK { return Err(ctx.error(st)); }
        let st = unsafe { ffi::acu_memcpy_h2d(ctx.raw, ptr, bytes.as_ptr() as *const c_void, bytes.len()) };
        if st != ffi::ACU_OK { return Err(ctx.error(st)); }
        Ok(Arc::new(Self { ctx: ctx.raw, ptr, len: bytes.len() }))
    }
    pub fn to_host(&self) -> Buffer {
        let mut v = vec![0u8; self.len];
        unsafe { ffi::acu_memcpy_d2h(self.ctx, v.as_mut_ptr() as *mut c_void, self.ptr, self.len) };
        Buffer::from_vec(v)
    }
}
impl Drop for DeviceBuffer { fn drop(&mut self) { unsafe { ffi::acu_free(self.ctx, self.ptr); } } }

/// Host array -> borrowed device view (uploads values + validity; bit offsets are preserved).
struct Uploaded { _values: Arc<DeviceBuffer>, _nulls: Option<Arc<DeviceBuffer>>, view: ffi::acu_array }

fn upload_primitive<T: ArrowPrimitiveType>(ctx: &Context, a: &PrimitiveArray<T>, is_scalar: bool) -> Result<Uploaded, ArrowError> {
    let values = DeviceBuffer::from_host(ctx, a.values().inner().as_slice())?;
    let (nulls, voff, nc, vptr) = match a.nulls() {
        Some(n) => { let b = DeviceBuffer::from_host(ctx, n.buffer().as_slice())?; let p = b.ptr as *const u8; (Some(b), n.offset() as i64, n.null_count() as i64, p) }
        None => (None, 0, 0, std::ptr::null()),
    };
    let view = ffi::acu_array { values: values.ptr, values_offset: 0, validity: vptr, validity_offset: voff, len: a.len() as i64,
                                null_count: nc, is_scalar: is_scalar as i32, reserved: 0 };
    Ok(Uploaded { _values: values, _nulls: nulls, view })
}

/// `arrow::compute::kernels::numeric::add` for primitive arrays of one native type.
/// (`sub`, `mul`, `div`, `rem`, `*_wrapping` differ only in the `op` code: include/arrow_cuda.h acu_arith_op.)
pub fn add<T: ArrowPrimitiveType>(ctx: &Context, dtype: i32, lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> {
    arithmetic_op::<T>(ctx, dtype, 1 /* ACU_ADD */, lhs, rhs)
}

fn arithmetic_op<T: ArrowPrimitiveType>(ctx: &Context, dtype: i32, op: i32, lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> {
    let (l, l_s) = lhs.get();
    let (r, r_s) = rhs.get();
    let l = l.as_any().downcast_ref::<PrimitiveArray<T>>().ok_or_else(|| ArrowError::InvalidArgumentError("type mismatch".into()))?;
    let r = r.as_any().downcast_ref::<PrimitiveArray<T>>().ok_or_else(|| ArrowError::InvalidArgumentError("type mismatch".into()))?;
    let (a, b) = (upload_primitive(ctx, l, l_s)?, upload_primitive(ctx, r, r_s)?);
    let n = if l_s && !r_s { r.len() } else { l.len() };
    let width = std::mem::size_of::<T::Native>();
    let out_values = DeviceBuffer::from_host(ctx, &vec![0u8; n.max(1) * width])?;
    let out_valid = DeviceBuffer::from_host(ctx, &vec![0u8; (n + 63) / 64 * 8 + 8])?;
    let mut out = ffi::acu_array_out { values: out_values.ptr, validity: out_valid.ptr as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 };
    let st = unsafe { ffi::acu_arith(ctx.raw, dtype, op, &a.view, &b.view, &mut out) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); }
    let values = ScalarBuffer::<T::Native>::new(out_values.to_host(), 0, out.len as usize);
    let nulls = (out.has_validity != 0).then(|| unsafe {
        NullBuffer::new_unchecked(BooleanBuffer::new(out_valid.to_host(), 0, out.len as usize), out.null_count as usize)
    });
    Ok(Arc::new(PrimitiveArray::<T>::new(values, nulls)))
}

/// `arrow::compute::filter` for primitive arrays: plan (FilterBuilder::new(..).build()) + compaction.
pub fn filter<T: ArrowPrimitiveType>(ctx: &Context, values: &PrimitiveArray<T>, predicate: &BooleanArray) -> Result<ArrayRef, ArrowError> {
    let pv = DeviceBuffer::from_host(ctx, predicate.values().inner().as_slice())?;
    let pn = predicate.nulls().map(|n| DeviceBuffer::from_host(ctx, n.buffer().as_slice())).transpose()?;
    let pred = ffi::acu_array { values: pv.ptr, values_offset: predicate.values().offset() as i64,
        validity: pn.as_ref().map_or(std::ptr::null(), |b| b.ptr as *const u8),
        validity_offset: predicate.nulls().map_or(0, |n| n.offset() as i64), len: predicate.len() as i64,
        null_count: predicate.null_count() as i64, is_scalar: 0, reserved: 0 };
    let mut plan = std::ptr::null_mut();
    let st = unsafe { ffi::acu_filter_plan_create(ctx.raw, &pred, &mut plan) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); }
    let count = unsafe { ffi::acu_filter_plan_count(plan) } as usize;
    let v = upload_primitive(ctx, values, false)?;
    let width = std::mem::size_of::<T::Native>();
    let out_values = DeviceBuffer::from_host(ctx, &vec![0u8; count.max(1) * width])?;
    let out_valid = DeviceBuffer::from_host(ctx, &vec![0u8; (count + 63) / 64 * 8 + 8])?;
    let mut out = ffi::acu_array_out { values: out_values.ptr, validity: out_valid.ptr as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 };
    let st = unsafe { ffi::acu_filter_primitive(ctx.raw, plan, width as i32, &v.view, &mut out) };
    unsafe { ffi::acu_filter_plan_destroy(ctx.raw, plan) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); }
    let vals = ScalarBuffer::<T::Native>::new(out_values.to_host(), 0, out.len as usize);
    let nulls = (out.has_validity != 0).then(|| unsafe {
        NullBuffer::new_unchecked(BooleanBuffer::new(out_valid.to_host(), 0, out.len as usize), out.null_count as usize)
    });
    let _ = DataType::Null;
    Ok(Arc::new(PrimitiveArray::<T>::new(vals, nulls).with_data_type(values.data_type().clone())))
}
/// Shared tail of every call that returns a primitive array: download values (+ validity when the
/// result carries a NullBuffer) and re-attach the logical DataType.
fn finish_primitive<T: ArrowPrimitiveType>(out: &ffi::acu_array_out, values: &DeviceBuffer, valid: &DeviceBuffer, data_type: &DataType) -> ArrayRef {
    let vals = ScalarBuffer::<T::Native>::new(values.to_host(), 0, out.len as usize);
    let nulls = (out.has_validity != 0).then(|| unsafe {
        NullBuffer::new_unchecked(BooleanBuffer::new(valid.to_host(), 0, out.len as usize), out.null_count as usize)
    });
    Arc::new(PrimitiveArray::<T>::new(vals, nulls).with_data_type(data_type.clone()))
}

/// `arrow::compute::take(values, indices, options)` (arrow-select/src/take.rs:89-105) for primitive values and any
/// integer index type `I`; `index_dtype` is the acu_dtype code of `I` (include/arrow_cuda.h).
pub fn take<T: ArrowPrimitiveType, I: ArrowPrimitiveType>(ctx: &Context, values: &PrimitiveArray<T>, indices: &PrimitiveArray<I>,
                                                          index_dtype: i32, check_bounds: bool) -> Result<ArrayRef, ArrowError> {
    let (v, ix) = (upload_primitive(ctx, values, false)?, upload_primitive(ctx, indices, false)?);
    let (m, width) = (indices.len(), std::mem::size_of::<T::Native>());
    let out_values = DeviceBuffer::from_host(ctx, &vec![0u8; m.max(1) * width])?;
    let out_valid = DeviceBuffer::from_host(ctx, &vec![0u8; (m + 63) / 64 * 8 + 8])?;
    let mut out = ffi::acu_array_out { values: out_values.ptr, validity: out_valid.ptr as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 };
    let st = unsafe { ffi::acu_take_primitive(ctx.raw, width as i32, &v.view, &ix.view, index_dtype, check_bounds as i32, &mut out) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); } // ACU_ERR_PANIC_OUT_OF_BOUNDS panics inside error(), like take.rs:447
    Ok(finish_primitive::<T>(&out, &out_values, &out_valid, values.data_type()))
}

/// `arrow::compute::kernels::cmp::{eq, neq, lt, lt_eq, gt, gt_eq, distinct, not_distinct}` (arrow-ord/src/cmp.rs:79-202):
/// `op` is the acu_cmp_op code; floats compare by IEEE totalOrder like the reference.
pub fn compare<T: ArrowPrimitiveType>(ctx: &Context, dtype: i32, op: i32, lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> {
    let (l, l_s) = lhs.get();
    let (r, r_s) = rhs.get();
    let l = l.as_any().downcast_ref::<PrimitiveArray<T>>().ok_or_else(|| ArrowError::InvalidArgumentError("type mismatch".into()))?;
    let r = r.as_any().downcast_ref::<PrimitiveArray<T>>().ok_or_else(|| ArrowError::InvalidArgumentError("type mismatch".into()))?;
    let (a, b) = (upload_primitive(ctx, l, l_s)?, upload_primitive(ctx, r, r_s)?);
    let n = if l_s { r.len() } else { l.len() };
    let bytes = (n.max(1) + 63) / 64 * 8 + 8;
    let (out_bits, out_valid) = (DeviceBuffer::from_host(ctx, &vec![0u8; bytes])?, DeviceBuffer::from_host(ctx, &vec![0u8; bytes])?);
    let mut out = ffi::acu_array_out { values: out_bits.ptr, validity: out_valid.ptr as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 };
    let st = unsafe { ffi::acu_cmp(ctx.raw, dtype, op, &a.view, &b.view, &mut out) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); }
    let values = BooleanBuffer::new(out_bits.to_host(), 0, out.len as usize);
    let nulls = (out.has_validity != 0).then(|| unsafe {
        NullBuffer::new_unchecked(BooleanBuffer::new(out_valid.to_host(), 0, out.len as usize), out.null_count as usize)
    });
    Ok(BooleanArray::new(values, nulls))
}

/// `arrow::compute::{sum, min, max}` (arrow-arith/src/aggregate.rs:943,1012,1027): `None` iff no valid row.
pub fn aggregate<T: ArrowPrimitiveType>(ctx: &Context, dtype: i32, op: i32, array: &PrimitiveArray<T>) -> Result<Option<T::Native>, ArrowError>
where T::Native: Copy {
    let a = upload_primitive(ctx, array, false)?;
    let (mut bits, mut valid) = (0u64, 0i64);
    let st = unsafe { ffi::acu_aggregate(ctx.raw, dtype, op, &a.view, &mut bits, &mut valid) };
    if st != ffi::ACU_OK { return Err(ctx.error(st)); }
    if valid == 0 { return Ok(None); }
    // the result is the native value's bit pattern, zero-extended to 64 bits (little endian)
    Ok(Some(unsafe { std::ptr::read_unaligned(&bits as *const u64 as *const T::Native) }))
}
// cast and the RecordBatch-level calls (acu_cast_numeric, acu_filter_record_batch, acu_take_record_batch) follow the
// same pattern; INTEGRATION.md §2 has the full mapping table. A production shim keeps arrays in `DeviceBuffer`s between
// calls instead of uploading / downloading around every kernel as these reference-shaped wrappers do.
