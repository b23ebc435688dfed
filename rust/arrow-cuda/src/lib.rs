//! `arrow-cuda`: drop-in for the `arrow::compute` hot path, backed by hand-written sm_100a kernels
//! (libarrow_cuda.so, C ABI in include/arrow_cuda.h).
//!
//! SOURCE ONLY in this repository — there is no Rust toolchain in the build image, so this crate is written to be correct by
//! inspection; the compiled, GPU-tested equivalents of this layer are `arrow-rs_b200/host/arrow_cuda.hpp` (C++) and
//! `arrow-rs_b200/acu` (Python). A caller switches by changing `use arrow::compute::...` to `use arrow_cuda::compute::...`:
//! every public function below has the reference's name, argument types and result type
//! (arrow/src/compute/mod.rs:20-40, arrow/src/compute/kernels.rs:20-34):
//!
//!   compute::{filter, filter_record_batch, FilterBuilder, FilterPredicate}   arrow-select/src/filter.rs:201-533
//!   compute::{take, take_arrays, take_record_batch, TakeOptions}             arrow-select/src/take.rs:89-164,1123
//!   compute::kernels::numeric::{add .. rem, neg, neg_wrapping}               arrow-arith/src/numeric.rs:36-186
//!   compute::kernels::cmp::{eq .. not_distinct}                              arrow-ord/src/cmp.rs:79-202
//!   compute::kernels::boolean::{and .. is_not_null}                          arrow-arith/src/boolean.rs:60-354
//!   compute::{cast, cast_with_options, CastOptions}                          arrow-cast/src/cast/mod.rs:347,790
//!   compute::{sum, min, max, sum_checked}                                    arrow-arith/src/aggregate.rs:897-1027
//!   compute::{nullif, zip, concat, concat_batches}                           arrow-select/src/{nullif,zip,concat}.rs
//!
//! The wrappers are reference-shaped: host `ArrayRef` in, host `ArrayRef` out, upload / download around every call. A
//! production integration keeps columns in [`DeviceBuffer`]s between calls (what the C++ mirror does); the device-resident
//! building blocks are public for that purpose ([`DeviceArray`], [`Context`]).
pub mod error;
pub mod ffi;

use arrow_array::{make_array, Array, ArrayRef, ArrowPrimitiveType, BooleanArray, Datum, PrimitiveArray, RecordBatch};
use arrow_buffer::{BooleanBuffer, Buffer, MutableBuffer, NullBuffer};
use arrow_data::ArrayData;
use arrow_schema::{ArrowError, DataType, SchemaRef};
use std::cell::RefCell;
use std::os::raw::c_void;
use std::rc::Rc;
use std::sync::Arc;

// ---------------------------------------------------------------------------------------------------------------------
// Context + DeviceBuffer
// ---------------------------------------------------------------------------------------------------------------------
struct ContextInner { raw: *mut ffi::acu_ctx }
impl Drop for ContextInner { fn drop(&mut self) { unsafe { ffi::acu_ctx_destroy(self.raw) } } }

/// One device + one stream (`acu_ctx`). The reference kernels are pure functions; the context is the implicit "where does
/// this run". `Rc`: a ctx must not be used from two threads at once (include/arrow_cuda.h), so it is neither Send nor Sync;
/// every thread gets its own default context.
#[derive(Clone)]
pub struct Context { inner: Rc<ContextInner> }

thread_local! { static DEFAULT: RefCell<Option<Context>> = RefCell::new(None); }

impl Context {
    pub fn new(device: i32) -> Result<Self, ArrowError> {
        let mut raw = std::ptr::null_mut();
        match unsafe { ffi::acu_ctx_create(device, &mut raw) } {
            ffi::ACU_OK => Ok(Self { inner: Rc::new(ContextInner { raw }) }),
            st => Err(ArrowError::ExternalError(format!("acu_ctx_create failed ({st}): no CUDA device, and there is no CPU fallback").into())),
        }
    }
    /// The calling thread's default context (device `ARROW_CUDA_DEVICE`, default 0), created on first use — what the
    /// reference-shaped free functions run on.
    pub fn current() -> Result<Self, ArrowError> {
        DEFAULT.with(|d| {
            if d.borrow().is_none() {
                let dev = std::env::var("ARROW_CUDA_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
                *d.borrow_mut() = Some(Context::new(dev)?);
            }
            Ok(d.borrow().as_ref().unwrap().clone())
        })
    }
    pub(crate) fn raw(&self) -> *mut ffi::acu_ctx { self.inner.raw }
    pub(crate) fn check(&self, st: ffi::acu_status) -> Result<(), ArrowError> {
        if st == ffi::ACU_OK { return Ok(()); }
        Err(error::from_detail(st, unsafe { &*ffi::acu_last_error(self.inner.raw) }))
    }
}

/// DeviceBuffer: `arrow_buffer::Buffer { data: Arc<Bytes>, ptr, length }` (arrow-buffer/src/buffer/immutable.rs:83-96) with
/// the bytes in HBM. It keeps its context alive (the `Rc<ContextInner>`), so it can never be freed on a destroyed ctx.
pub struct DeviceBuffer { ctx: Context, ptr: *mut c_void, len: usize }
impl DeviceBuffer {
    /// `len` bytes (+ 16 bytes of slack: kernels read whole aligned words), uninitialised.
    pub fn allocate(ctx: &Context, len: usize) -> Result<Self, ArrowError> {
        let mut ptr = std::ptr::null_mut();
        ctx.check(unsafe { ffi::acu_malloc(ctx.raw(), len + 16, &mut ptr) })?;
        Ok(Self { ctx: ctx.clone(), ptr, len })
    }
    pub fn from_host(ctx: &Context, bytes: &[u8]) -> Result<Self, ArrowError> {
        let b = Self::allocate(ctx, bytes.len())?;
        if !bytes.is_empty() {
            ctx.check(unsafe { ffi::acu_memcpy_h2d(ctx.raw(), b.ptr, bytes.as_ptr() as *const c_void, bytes.len()) })?;
        }
        Ok(b)
    }
    /// The first `len` bytes as an arrow `Buffer` (128-byte aligned `MutableBuffer`, so any native type can view it).
    pub fn to_host(&self, len: usize) -> Result<Buffer, ArrowError> {
        let len = len.min(self.len);
        let mut m = MutableBuffer::from_len_zeroed(len);
        if len > 0 {
            self.ctx.check(unsafe { ffi::acu_memcpy_d2h(self.ctx.raw(), m.as_mut_ptr() as *mut c_void, self.ptr, len) })?;
        }
        Ok(m.into())
    }
    pub fn as_ptr(&self) -> *mut c_void { self.ptr }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
}
impl Drop for DeviceBuffer { fn drop(&mut self) { unsafe { ffi::acu_free(self.ctx.raw(), self.ptr); } } }

fn bitmap_bytes(rows: usize) -> usize { (rows + 63) / 64 * 8 }

// ---------------------------------------------------------------------------------------------------------------------
// Host array -> device view (generic over the array kinds of the hot path, via ArrayData's buffer layout)
// ---------------------------------------------------------------------------------------------------------------------
/// acu_dtype code of a numeric DataType (the order of include/arrow_cuda.h `acu_dtype`).
fn dtype_code(t: &DataType) -> Option<i32> {
    Some(match t {
        DataType::Int8 => ffi::ACU_I8, DataType::Int16 => ffi::ACU_I16, DataType::Int32 => ffi::ACU_I32, DataType::Int64 => ffi::ACU_I64,
        DataType::UInt8 => ffi::ACU_U8, DataType::UInt16 => ffi::ACU_U16, DataType::UInt32 => ffi::ACU_U32, DataType::UInt64 => ffi::ACU_U64,
        DataType::Float32 => ffi::ACU_F32, DataType::Float64 => ffi::ACU_F64,
        _ => return None,
    })
}

enum Kind { Primitive(usize), Boolean, Bytes(usize) }
fn kind_of(t: &DataType) -> Result<Kind, ArrowError> {
    match t {
        DataType::Boolean => Ok(Kind::Boolean),
        DataType::Utf8 | DataType::Binary => Ok(Kind::Bytes(4)),
        DataType::LargeUtf8 | DataType::LargeBinary => Ok(Kind::Bytes(8)),
        // filter / take are type-agnostic copies: every fixed-width primitive goes by element width
        t => t.primitive_width().map(Kind::Primitive).ok_or_else(|| ArrowError::NotYetImplemented(format!("arrow-cuda: data type {t}"))),
    }
}

/// A host array uploaded to HBM: owns the device copies, `column` is the borrowed view the C ABI takes.
pub struct DeviceArray {
    _bufs: Vec<DeviceBuffer>,
    pub column: ffi::acu_column,
    data_bytes: usize,
}

impl DeviceArray {
    pub fn upload(ctx: &Context, array: &dyn Array, is_scalar: bool) -> Result<Self, ArrowError> {
        let d: ArrayData = array.to_data();
        let kind = kind_of(d.data_type())?;
        let mut bufs = Vec::new();
        let (validity, validity_offset, null_count) = match d.nulls() {
            Some(n) => {
                let b = DeviceBuffer::from_host(ctx, n.buffer().as_slice())?;
                let p = b.as_ptr() as *const u8;
                bufs.push(b);
                (p, n.offset() as i64, n.null_count() as i64)
            }
            None => (std::ptr::null(), 0, 0),
        };
        let mut a = ffi::acu_array { values: std::ptr::null(), values_offset: 0, validity, validity_offset, len: d.len() as i64,
                                     null_count, is_scalar: is_scalar as i32, reserved: 0 };
        let mut col = ffi::acu_column { kind: ffi::ACU_COL_PRIMITIVE, width: 0, array: a, data: std::ptr::null() };
        let mut data_bytes = 0;
        match kind {
            Kind::Primitive(w) => {
                let bytes = &d.buffers()[0].as_slice()[d.offset() * w..(d.offset() + d.len()) * w];
                let b = DeviceBuffer::from_host(ctx, bytes)?;
                a.values = b.as_ptr();
                bufs.push(b);
                col.width = w as i32;
            }
            Kind::Boolean => {
                let b = DeviceBuffer::from_host(ctx, d.buffers()[0].as_slice())?;
                a.values = b.as_ptr();
                a.values_offset = d.offset() as i64;
                bufs.push(b);
                col.kind = ffi::ACU_COL_BOOLEAN;
            }
            Kind::Bytes(ob) => {
                let offs = &d.buffers()[0].as_slice()[d.offset() * ob..(d.offset() + d.len() + 1) * ob];
                let o = DeviceBuffer::from_host(ctx, offs)?;
                let v = DeviceBuffer::from_host(ctx, d.buffers()[1].as_slice())?;
                a.values = o.as_ptr();
                col.data = v.as_ptr() as *const u8;
                data_bytes = v.len();
                bufs.push(o);
                bufs.push(v);
                col.kind = ffi::ACU_COL_BYTES;
                col.width = ob as i32;
            }
        }
        col.array = a;
        Ok(Self { _bufs: bufs, column: col, data_bytes })
    }
    fn view(&self) -> &ffi::acu_array { &self.column.array }
}

/// Caller-owned output of one column + the download back into an `ArrayRef` of `data_type`.
struct ColumnOut { values: DeviceBuffer, validity: DeviceBuffer, data: Option<DeviceBuffer>, out: ffi::acu_column_out }
impl ColumnOut {
    fn new(ctx: &Context, data_type: &DataType, rows: usize, data_capacity: usize) -> Result<Self, ArrowError> {
        let vbytes = match kind_of(data_type)? {
            Kind::Primitive(w) => rows.max(1) * w,
            Kind::Boolean => bitmap_bytes(rows.max(1)),
            Kind::Bytes(ob) => (rows + 1) * ob,
        };
        let values = DeviceBuffer::allocate(ctx, vbytes)?;
        let validity = DeviceBuffer::allocate(ctx, bitmap_bytes(rows.max(1)))?;
        let data = match kind_of(data_type)? { Kind::Bytes(_) => Some(DeviceBuffer::allocate(ctx, data_capacity)?), _ => None };
        let out = ffi::acu_column_out {
            array: ffi::acu_array_out { values: values.as_ptr(), validity: validity.as_ptr() as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 },
            data: data.as_ref().map_or(std::ptr::null_mut(), |d| d.as_ptr() as *mut u8),
            data_capacity: data_capacity as i64,
            data_len: 0,
        };
        Ok(Self { values, validity, data, out })
    }
    fn array_out(&mut self) -> *mut ffi::acu_array_out { &mut self.out.array }
    fn finish(&self, data_type: &DataType) -> Result<ArrayRef, ArrowError> {
        let o = &self.out.array;
        let len = o.len as usize;
        let nulls = if o.has_validity != 0 {
            let bits = BooleanBuffer::new(self.validity.to_host(bitmap_bytes(len))?, 0, len);
            Some(unsafe { NullBuffer::new_unchecked(bits, o.null_count as usize) })
        } else { None };
        let mut b = ArrayData::builder(data_type.clone()).len(len).nulls(nulls);
        match kind_of(data_type)? {
            Kind::Primitive(w) => { b = b.add_buffer(self.values.to_host(len * w)?); }
            Kind::Boolean => { b = b.add_buffer(self.values.to_host(bitmap_bytes(len))?); }
            Kind::Bytes(ob) => {
                b = b.add_buffer(self.values.to_host((len + 1) * ob)?);
                b = b.add_buffer(self.data.as_ref().unwrap().to_host(self.out.data_len as usize)?);
            }
        }
        // the kernels produce valid Arrow buffers (bit-exact with the reference, tests/): no second validation pass
        Ok(make_array(unsafe { b.build_unchecked() }))
    }
}

fn numeric_dtype(ctx_what: &str, t: &DataType) -> Result<i32, ArrowError> {
    dtype_code(t).ok_or_else(|| ArrowError::InvalidArgumentError(format!("Invalid {ctx_what} operation: {t}")))
}

// ---------------------------------------------------------------------------------------------------------------------
// compute — the reference's public names and signatures
// ---------------------------------------------------------------------------------------------------------------------
pub mod compute {
    use super::*;

    // ---- filter (arrow-select/src/filter.rs) ------------------------------------------------------------------------
    /// RAII owner of the device-resident plan (`FilterPredicate`, filter.rs:442-533): freed on every path.
    struct Plan { ctx: Context, raw: *mut ffi::acu_filter_plan }
    impl Drop for Plan { fn drop(&mut self) { unsafe { ffi::acu_filter_plan_destroy(self.ctx.raw(), self.raw) } } }

    /// `FilterBuilder` (filter.rs:254-324). `optimize()` is a no-op here: the plan is always the "optimized" form.
    pub struct FilterBuilder { predicate: BooleanArray }
    impl FilterBuilder {
        pub fn new(filter: &BooleanArray) -> Self { Self { predicate: filter.clone() } }
        pub fn optimize(self) -> Self { self }
        pub fn build(self) -> Result<FilterPredicate, ArrowError> {
            let ctx = Context::current()?;
            let p = DeviceArray::upload(&ctx, &self.predicate, false)?;
            let mut raw = std::ptr::null_mut();
            ctx.check(unsafe { ffi::acu_filter_plan_create(ctx.raw(), p.view(), &mut raw) })?;
            Ok(FilterPredicate { plan: Plan { ctx, raw } })
        }
    }

    /// `FilterPredicate` (filter.rs:442-533): one scan of the predicate, reused for any number of arrays / batches.
    pub struct FilterPredicate { plan: Plan }
    impl FilterPredicate {
        /// Number of rows selected (FilterPredicate::count).
        pub fn count(&self) -> usize { unsafe { ffi::acu_filter_plan_count(self.plan.raw) as usize } }
        pub fn filter(&self, values: &dyn Array) -> Result<ArrayRef, ArrowError> {
            let ctx = &self.plan.ctx;
            let v = DeviceArray::upload(ctx, values, false)?;
            let mut out = ColumnOut::new(ctx, values.data_type(), self.count(), v.data_bytes)?;
            let st = match kind_of(values.data_type())? {
                Kind::Primitive(w) => unsafe { ffi::acu_filter_primitive(ctx.raw(), self.plan.raw, w as i32, v.view(), out.array_out()) },
                Kind::Boolean => unsafe { ffi::acu_filter_boolean(ctx.raw(), self.plan.raw, v.view(), out.array_out()) },
                Kind::Bytes(ob) => unsafe {
                    ffi::acu_filter_bytes(ctx.raw(), self.plan.raw, ob as i32, v.view().values, v.column.data, v.view(), out.out.array.values,
                                          out.out.data, out.out.data_capacity, &mut out.out.data_len, &mut out.out.array)
                },
            };
            ctx.check(st)?;
            out.finish(values.data_type())
        }
        /// filter.rs:459-478: every column with the same plan, one stream synchronisation per 64 columns.
        pub fn filter_record_batch(&self, record_batch: &RecordBatch) -> Result<RecordBatch, ArrowError> {
            let ctx = &self.plan.ctx;
            let n = record_batch.num_columns();
            let ups = record_batch.columns().iter().map(|c| DeviceArray::upload(ctx, c.as_ref(), false)).collect::<Result<Vec<_>, _>>()?;
            let mut outs = record_batch.columns().iter().zip(&ups).map(|(c, u)| ColumnOut::new(ctx, c.data_type(), self.count(), u.data_bytes))
                .collect::<Result<Vec<_>, _>>()?;
            for first in (0..n).step_by(ffi::ACU_MAX_BATCH_COLUMNS) {
                let last = (first + ffi::ACU_MAX_BATCH_COLUMNS).min(n);
                let cols: Vec<ffi::acu_column> = ups[first..last].iter().map(|u| u.column).collect();
                let mut raw: Vec<ffi::acu_column_out> = outs[first..last].iter().map(|o| o.out).collect();
                ctx.check(unsafe { ffi::acu_filter_record_batch(ctx.raw(), self.plan.raw, (last - first) as i32, cols.as_ptr(), raw.as_mut_ptr()) })?;
                for (o, r) in outs[first..last].iter_mut().zip(raw) { o.out = r; }
            }
            let cols = record_batch.columns().iter().zip(&outs).map(|(c, o)| o.finish(c.data_type())).collect::<Result<Vec<_>, _>>()?;
            RecordBatch::try_new(record_batch.schema(), cols)
        }
    }

    /// `arrow::compute::filter` (filter.rs:201-213).
    pub fn filter(values: &dyn Array, predicate: &BooleanArray) -> Result<ArrayRef, ArrowError> {
        FilterBuilder::new(predicate).build()?.filter(values)
    }
    /// `arrow::compute::filter_record_batch` (filter.rs:225-244).
    pub fn filter_record_batch(record_batch: &RecordBatch, predicate: &BooleanArray) -> Result<RecordBatch, ArrowError> {
        FilterBuilder::new(predicate).build()?.filter_record_batch(record_batch)
    }

    // ---- take (arrow-select/src/take.rs) ------------------------------------------------------------------------------
    /// take.rs:388-394
    #[derive(Clone, Debug, Default)]
    pub struct TakeOptions { pub check_bounds: bool }

    fn index_dtype(indices: &dyn Array) -> Result<i32, ArrowError> {
        match indices.data_type() {  // take.rs:96-104 downcast_integer_array!
            DataType::Float32 | DataType::Float64 => None,
            t => dtype_code(t),
        }.ok_or_else(|| ArrowError::InvalidArgumentError(format!("Take only supported for integers, got {:?}", indices.data_type())))
    }

    /// `arrow::compute::take` (take.rs:89-105).
    pub fn take(values: &dyn Array, indices: &dyn Array, options: Option<TakeOptions>) -> Result<ArrayRef, ArrowError> {
        let ctx = Context::current()?;
        let idt = index_dtype(indices)?;
        let check = options.unwrap_or_default().check_bounds as i32;
        let (v, ix) = (DeviceArray::upload(&ctx, values, false)?, DeviceArray::upload(&ctx, indices, false)?);
        let m = indices.len();
        let st;
        let mut out;
        match kind_of(values.data_type())? {
            Kind::Primitive(w) => {
                out = ColumnOut::new(&ctx, values.data_type(), m, 0)?;
                st = unsafe { ffi::acu_take_primitive(ctx.raw(), w as i32, v.view(), ix.view(), idt, check, out.array_out()) };
            }
            Kind::Boolean => {
                out = ColumnOut::new(&ctx, values.data_type(), m, 0)?;
                st = unsafe { ffi::acu_take_boolean(ctx.raw(), v.view(), ix.view(), idt, check, out.array_out()) };
            }
            Kind::Bytes(ob) => {
                // two-phase: offsets + required bytes first, then the copy (take_bytes computes the capacity first too, take.rs:520-523)
                let mut probe = ColumnOut::new(&ctx, values.data_type(), m, 0)?;
                let mut need = 0i64;
                ctx.check(unsafe {
                    ffi::acu_take_bytes(ctx.raw(), ob as i32, v.view().values, v.column.data, v.view(), ix.view(), idt, check, probe.out.array.values,
                                        std::ptr::null_mut(), 0, &mut need, &mut probe.out.array)
                })?;
                out = ColumnOut::new(&ctx, values.data_type(), m, need as usize)?;
                st = unsafe {
                    ffi::acu_take_bytes(ctx.raw(), ob as i32, v.view().values, v.column.data, v.view(), ix.view(), idt, check, out.out.array.values,
                                        out.out.data, out.out.data_capacity, &mut out.out.data_len, &mut out.out.array)
                };
            }
        }
        ctx.check(st)?; // ACU_ERR_PANIC_OUT_OF_BOUNDS panics inside, like take.rs:447
        out.finish(values.data_type())
    }
    /// `arrow::compute::take_arrays` (take.rs:155-164).
    pub fn take_arrays(arrays: &[ArrayRef], indices: &dyn Array, options: Option<TakeOptions>) -> Result<Vec<ArrayRef>, ArrowError> {
        arrays.iter().map(|a| take(a.as_ref(), indices, options.clone())).collect()
    }
    /// `arrow::compute::take_record_batch` (take.rs:1123-1133).
    pub fn take_record_batch(record_batch: &RecordBatch, indices: &dyn Array) -> Result<RecordBatch, ArrowError> {
        let cols = take_arrays(record_batch.columns(), indices, None)?;
        RecordBatch::try_new(record_batch.schema(), cols)
    }

    // ---- cast (arrow-cast/src/cast/mod.rs) -----------------------------------------------------------------------------
    /// mod.rs:96-111 (format options are irrelevant to numeric casts)
    #[derive(Clone, Debug)]
    pub struct CastOptions { pub safe: bool }
    impl Default for CastOptions { fn default() -> Self { Self { safe: true } } }

    /// `arrow::compute::cast` (mod.rs:347-349).
    pub fn cast(array: &dyn Array, to_type: &DataType) -> Result<ArrayRef, ArrowError> { cast_with_options(array, to_type, &CastOptions::default()) }
    /// `arrow::compute::cast_with_options` (mod.rs:790): numeric -> numeric on the device; Dictionary<_, Utf8> -> Utf8 through take
    /// (arrow-cast/src/cast/dictionary.rs:310-317); any other pair is outside the hot path (out of scope).
    pub fn cast_with_options(array: &dyn Array, to_type: &DataType, options: &CastOptions) -> Result<ArrayRef, ArrowError> {
        if array.data_type() == to_type { return Ok(make_array(array.to_data())); }
        if let (DataType::Dictionary(_, v), DataType::Utf8) = (array.data_type(), to_type) {
            if **v == DataType::Utf8 {
                let d = array.to_data();
                let values = make_array(d.child_data()[0].clone());
                let keys = make_array(d.clone().into_builder().data_type(match array.data_type() { DataType::Dictionary(k, _) => (**k).clone(), _ => unreachable!() })
                    .child_data(vec![]).build()?);
                return take(values.as_ref(), keys.as_ref(), None);
            }
        }
        let (from, to) = match (dtype_code(array.data_type()), dtype_code(to_type)) {
            (Some(f), Some(t)) => (f, t),
            _ => return Err(ArrowError::CastError(format!("Casting from {} to {} not supported", array.data_type(), to_type))),
        };
        let ctx = Context::current()?;
        let a = DeviceArray::upload(&ctx, array, false)?;
        let mut out = ColumnOut::new(&ctx, to_type, array.len(), 0)?;
        ctx.check(unsafe { ffi::acu_cast_numeric(ctx.raw(), from, to, options.safe as i32, a.view(), out.array_out()) })?;
        out.finish(to_type)
    }

    // ---- aggregate (arrow-arith/src/aggregate.rs) ------------------------------------------------------------------------
    fn aggregate<T: ArrowPrimitiveType>(op: i32, array: &PrimitiveArray<T>) -> Option<T::Native> {
        let dtype = dtype_code(&T::DATA_TYPE)?; // derived from T: a caller cannot pass a mismatching code
        let ctx = Context::current().ok()?;
        let a = DeviceArray::upload(&ctx, array, false).ok()?;
        let (mut bits, mut valid) = (0u64, 0i64);
        ctx.check(unsafe { ffi::acu_aggregate(ctx.raw(), dtype, op, a.view(), &mut bits, &mut valid) }).ok()?;
        if valid == 0 { return None; }
        Some(unsafe { std::ptr::read_unaligned(&bits as *const u64 as *const T::Native) }) // native bit pattern, zero-extended, little endian
    }
    /// aggregate.rs:943 / :1012 / :1027 — `None` iff no valid row.
    pub fn sum<T: ArrowPrimitiveType>(array: &PrimitiveArray<T>) -> Option<T::Native> { aggregate(ffi::ACU_SUM, array) }
    pub fn min<T: ArrowPrimitiveType>(array: &PrimitiveArray<T>) -> Option<T::Native> { aggregate(ffi::ACU_MIN, array) }
    pub fn max<T: ArrowPrimitiveType>(array: &PrimitiveArray<T>) -> Option<T::Native> { aggregate(ffi::ACU_MAX, array) }
    /// aggregate.rs:897-937
    pub fn sum_checked<T: ArrowPrimitiveType>(array: &PrimitiveArray<T>) -> Result<Option<T::Native>, ArrowError> {
        let dtype = numeric_dtype("arithmetic", &T::DATA_TYPE)?;
        let ctx = Context::current()?;
        let a = DeviceArray::upload(&ctx, array, false)?;
        let (mut bits, mut valid) = (0u64, 0i64);
        ctx.check(unsafe { ffi::acu_sum_checked(ctx.raw(), dtype, a.view(), &mut bits, &mut valid) })?;
        Ok((valid != 0).then(|| unsafe { std::ptr::read_unaligned(&bits as *const u64 as *const T::Native) }))
    }

    // ---- nullif / zip / concat (arrow-select) ----------------------------------------------------------------------------
    /// `arrow::compute::nullif` (nullif.rs:44): values shared, validity &= !(right is Some(true)).
    pub fn nullif(left: &dyn Array, right: &BooleanArray) -> Result<ArrayRef, ArrowError> {
        let ctx = Context::current()?;
        let (l, r) = (DeviceArray::upload(&ctx, left, false)?, DeviceArray::upload(&ctx, right, false)?);
        let validity = DeviceBuffer::allocate(&ctx, bitmap_bytes(left.len().max(1)))?;
        let mut out = ffi::acu_array_out { values: std::ptr::null_mut(), validity: validity.as_ptr() as *mut u8, len: 0, null_count: 0, has_validity: 0, reserved: 0 };
        ctx.check(unsafe { ffi::acu_nullif(ctx.raw(), l.view(), r.view(), &mut out) })?;
        if left.is_empty() { return Ok(make_array(left.to_data())); }
        let nulls = if out.has_validity != 0 {
            Some(unsafe { NullBuffer::new_unchecked(BooleanBuffer::new(validity.to_host(bitmap_bytes(left.len()))?, 0, left.len()), out.null_count as usize) })
        } else { None };
        // only the null mask changes; the host value buffers are shared (nullif.rs:107-112). The slice is normalised to
        // offset 0 first because the new mask has bit offset 0.
        let d = left.to_data();
        let d = if d.offset() != 0 {
            let mut m = arrow_data::transform::MutableArrayData::new(vec![&d], false, d.len());
            m.extend(0, 0, d.len());
            m.freeze()
        } else { d };
        Ok(make_array(unsafe { d.into_builder().nulls(nulls).build_unchecked() }))
    }

    /// `arrow::compute::zip` (zip.rs:99) for primitive arrays / scalars.
    pub fn zip(mask: &BooleanArray, truthy: &dyn Datum, falsy: &dyn Datum) -> Result<ArrayRef, ArrowError> {
        let (t, t_s) = truthy.get();
        let (f, f_s) = falsy.get();
        if t.data_type() != f.data_type() { return Err(ArrowError::InvalidArgumentError("arguments need to have the same data type".into())); }
        let w = t.data_type().primitive_width().ok_or_else(|| ArrowError::NotYetImplemented(format!("arrow-cuda zip: {}", t.data_type())))?;
        let ctx = Context::current()?;
        let (m, tv, fv) = (DeviceArray::upload(&ctx, mask, false)?, DeviceArray::upload(&ctx, t, t_s)?, DeviceArray::upload(&ctx, f, f_s)?);
        let mut out = ColumnOut::new(&ctx, t.data_type(), mask.len(), 0)?;
        ctx.check(unsafe { ffi::acu_zip(ctx.raw(), w as i32, m.view(), tv.view(), fv.view(), out.array_out()) })?;
        out.finish(t.data_type())
    }

    /// `arrow::compute::concat` (concat.rs:495).
    pub fn concat(arrays: &[&dyn Array]) -> Result<ArrayRef, ArrowError> {
        if arrays.is_empty() { return Err(ArrowError::ComputeError("concat requires input of at least one array".into())); }
        if arrays.len() == 1 { return Ok(arrays[0].slice(0, arrays[0].len())); }
        let d = arrays[0].data_type();
        if arrays.iter().skip(1).any(|a| a.data_type() != d) { // the reference lists up to 10 distinct types (concat.rs:505-535)
            let mut seen: Vec<&DataType> = vec![d];
            let mut msg = format!("It is not possible to concatenate arrays of different data types ({d}");
            for a in arrays {
                if !seen.contains(&a.data_type()) {
                    seen.push(a.data_type());
                    if seen.len() == 11 { msg.push_str(", ..."); break; }
                    msg.push_str(", ");
                    msg.push_str(&a.data_type().to_string());
                }
            }
            msg.push_str(").");
            return Err(ArrowError::InvalidArgumentError(msg));
        }
        let ctx = Context::current()?;
        let ups = arrays.iter().map(|a| DeviceArray::upload(&ctx, *a, false)).collect::<Result<Vec<_>, _>>()?;
        let cols: Vec<ffi::acu_column> = ups.iter().map(|u| u.column).collect();
        let rows: usize = arrays.iter().map(|a| a.len()).sum();
        let mut out = ColumnOut::new(&ctx, d, rows, ups.iter().map(|u| u.data_bytes).sum())?;
        ctx.check(unsafe { ffi::acu_concat(ctx.raw(), cols.len() as i32, cols.as_ptr(), &mut out.out) })?;
        out.finish(d)
    }
    /// `arrow::compute::concat_batches` (concat.rs:607).
    pub fn concat_batches<'a>(schema: &SchemaRef, input_batches: impl IntoIterator<Item = &'a RecordBatch>) -> Result<RecordBatch, ArrowError> {
        let batches: Vec<&RecordBatch> = input_batches.into_iter().collect();
        if schema.fields().is_empty() {
            let rows = batches.iter().map(|b| b.num_rows()).sum();
            return RecordBatch::try_new_with_options(schema.clone(), vec![], &arrow_array::RecordBatchOptions::new().with_row_count(Some(rows)));
        }
        if batches.is_empty() { return Ok(RecordBatch::new_empty(schema.clone())); }
        let cols = (0..schema.fields().len())
            .map(|i| concat(&batches.iter().map(|b| b.column(i).as_ref()).collect::<Vec<_>>()))
            .collect::<Result<Vec<_>, _>>()?;
        RecordBatch::try_new(schema.clone(), cols)
    }

    // ---- kernels ------------------------------------------------------------------------------------------------------------
    pub mod kernels {
        use super::*;

        /// `arrow::compute::kernels::numeric` (arrow-arith/src/numeric.rs:36-186)
        pub mod numeric {
            use super::*;
            fn arithmetic_op(op: i32, sym: &str, lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> {
                let (l, l_s) = lhs.get();
                let (r, r_s) = rhs.get();
                let dtype = match (dtype_code(l.data_type()), l.data_type() == r.data_type()) { // numeric.rs:270-272
                    (Some(c), true) => c,
                    _ => return Err(ArrowError::InvalidArgumentError(format!("Invalid arithmetic operation: {} {sym} {}", l.data_type(), r.data_type()))),
                };
                let ctx = Context::current()?;
                let (a, b) = (DeviceArray::upload(&ctx, l, l_s)?, DeviceArray::upload(&ctx, r, r_s)?);
                let n = if l_s && !r_s { r.len() } else { l.len() };
                let mut out = ColumnOut::new(&ctx, l.data_type(), n, 0)?;
                ctx.check(unsafe { ffi::acu_arith(ctx.raw(), dtype, op, a.view(), b.view(), out.array_out()) })?;
                out.finish(l.data_type())
            }
            pub fn add(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_ADD, "+", lhs, rhs) }
            pub fn add_wrapping(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_ADD_WRAPPING, "+", lhs, rhs) }
            pub fn sub(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_SUB, "-", lhs, rhs) }
            pub fn sub_wrapping(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_SUB_WRAPPING, "-", lhs, rhs) }
            pub fn mul(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_MUL, "*", lhs, rhs) }
            pub fn mul_wrapping(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_MUL_WRAPPING, "*", lhs, rhs) }
            pub fn div(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_DIV, "/", lhs, rhs) }
            pub fn rem(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<ArrayRef, ArrowError> { arithmetic_op(ffi::ACU_REM, "%", lhs, rhs) }
            fn neg_op(array: &dyn Array, checked: i32) -> Result<ArrayRef, ArrowError> {
                let dtype = numeric_dtype("arithmetic", array.data_type())?;
                let ctx = Context::current()?;
                let a = DeviceArray::upload(&ctx, array, false)?;
                let mut out = ColumnOut::new(&ctx, array.data_type(), array.len(), 0)?;
                ctx.check(unsafe { ffi::acu_neg(ctx.raw(), dtype, checked, a.view(), out.array_out()) })?;
                out.finish(array.data_type())
            }
            pub fn neg(array: &dyn Array) -> Result<ArrayRef, ArrowError> { neg_op(array, 1) }
            pub fn neg_wrapping(array: &dyn Array) -> Result<ArrayRef, ArrowError> { neg_op(array, 0) }
        }

        /// `arrow::compute::kernels::cmp` (arrow-ord/src/cmp.rs:79-202)
        pub mod cmp {
            use super::*;
            fn compare_op(op: i32, sym: &str, lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> {
                let (l, l_s) = lhs.get();
                let (r, r_s) = rhs.get();
                if l.data_type() != r.data_type() { // cmp.rs:260-264
                    return Err(ArrowError::InvalidArgumentError(format!("Invalid comparison operation: {} {sym} {}", l.data_type(), r.data_type())));
                }
                let ctx = Context::current()?;
                let (a, b) = (DeviceArray::upload(&ctx, l, l_s)?, DeviceArray::upload(&ctx, r, r_s)?);
                let n = if l_s { r.len() } else { l.len() };
                let mut out = ColumnOut::new(&ctx, &DataType::Boolean, n, 0)?;
                let st = match kind_of(l.data_type())? {
                    Kind::Bytes(ob) => {
                        let (x, y) = (ffi::acu_bytes_array { offsets: a.view().values, data: a.column.data, nulls: *a.view() },
                                      ffi::acu_bytes_array { offsets: b.view().values, data: b.column.data, nulls: *b.view() });
                        unsafe { ffi::acu_cmp_bytes(ctx.raw(), ob as i32, op, &x, &y, out.array_out()) }
                    }
                    _ => {
                        let dtype = dtype_code(l.data_type())
                            .ok_or_else(|| ArrowError::InvalidArgumentError(format!("Invalid comparison operation: {} {sym} {}", l.data_type(), r.data_type())))?;
                        unsafe { ffi::acu_cmp(ctx.raw(), dtype, op, a.view(), b.view(), out.array_out()) }
                    }
                };
                ctx.check(st)?;
                let r = out.finish(&DataType::Boolean)?;
                Ok(r.as_any().downcast_ref::<BooleanArray>().unwrap().clone())
            }
            pub fn eq(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_EQ, "==", lhs, rhs) }
            pub fn neq(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_NEQ, "!=", lhs, rhs) }
            pub fn lt(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_LT, "<", lhs, rhs) }
            pub fn lt_eq(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_LT_EQ, "<=", lhs, rhs) }
            pub fn gt(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_GT, ">", lhs, rhs) }
            pub fn gt_eq(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_GT_EQ, ">=", lhs, rhs) }
            pub fn distinct(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_DISTINCT, "IS DISTINCT FROM", lhs, rhs) }
            pub fn not_distinct(lhs: &dyn Datum, rhs: &dyn Datum) -> Result<BooleanArray, ArrowError> { compare_op(ffi::ACU_NOT_DISTINCT, "IS NOT DISTINCT FROM", lhs, rhs) }
        }

        /// `arrow::compute::kernels::boolean` (arrow-arith/src/boolean.rs:60-354)
        pub mod boolean {
            use super::*;
            fn boolean_op(op: i32, a: &dyn Array, b: Option<&BooleanArray>) -> Result<BooleanArray, ArrowError> {
                let ctx = Context::current()?;
                let da = DeviceArray::upload(&ctx, a, false)?;
                let db = b.map(|b| DeviceArray::upload(&ctx, b, false)).transpose()?;
                let mut out = ColumnOut::new(&ctx, &DataType::Boolean, a.len(), 0)?;
                ctx.check(unsafe { ffi::acu_boolean(ctx.raw(), op, da.view(), db.as_ref().map_or(std::ptr::null(), |d| d.view() as *const _), out.array_out()) })?;
                Ok(out.finish(&DataType::Boolean)?.as_any().downcast_ref::<BooleanArray>().unwrap().clone())
            }
            pub fn and(left: &BooleanArray, right: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_AND, left, Some(right)) }
            pub fn or(left: &BooleanArray, right: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_OR, left, Some(right)) }
            pub fn and_not(left: &BooleanArray, right: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_AND_NOT, left, Some(right)) }
            pub fn and_kleene(left: &BooleanArray, right: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_AND_KLEENE, left, Some(right)) }
            pub fn or_kleene(left: &BooleanArray, right: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_OR_KLEENE, left, Some(right)) }
            pub fn not(left: &BooleanArray) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_NOT, left, None) }
            pub fn is_null(input: &dyn Array) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_IS_NULL, input, None) }
            pub fn is_not_null(input: &dyn Array) -> Result<BooleanArray, ArrowError> { boolean_op(ffi::ACU_BOOL_IS_NOT_NULL, input, None) }
        }
    }
}
