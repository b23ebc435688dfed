//! `acu_status` + `acu_error_detail` -> the `ArrowError` variant (and message) the reference returns.
//!
//! Every data-dependent failure of the hot path has a pinned `Display` text in the reference; the C ABI
//! rebuilds that text in `acu_error_detail::message` *including* the variant prefix ("Compute error: ..."), so the shim only
//! strips the prefix again when it re-wraps the message in the variant. The reference's out-of-bounds `take` panics
//! (arrow-select/src/take.rs:447,454) stay panics.
use crate::ffi;
use arrow_schema::ArrowError;
use std::ffi::CStr;

pub(crate) fn from_detail(st: ffi::acu_status, d: &ffi::acu_error_detail) -> ArrowError {
    let full = unsafe { CStr::from_ptr(d.message.as_ptr()) }.to_string_lossy().into_owned();
    let strip = |p: &str| full.strip_prefix(p).unwrap_or(&full).to_string();
    match st {
        ffi::ACU_ERR_INVALID_ARGUMENT => ArrowError::InvalidArgumentError(strip("Invalid argument error: ")),
        ffi::ACU_ERR_COMPUTE => ArrowError::ComputeError(strip("Compute error: ")),
        ffi::ACU_ERR_ARITHMETIC_OVERFLOW => ArrowError::ArithmeticOverflow(strip("Arithmetic overflow: ")),
        ffi::ACU_ERR_DIVIDE_BY_ZERO => ArrowError::DivideByZero,
        ffi::ACU_ERR_OFFSET_OVERFLOW => ArrowError::OffsetOverflowError(d.len as usize),
        ffi::ACU_ERR_CAST => ArrowError::CastError(strip("Cast error: ")),
        ffi::ACU_ERR_NOT_YET_IMPLEMENTED => ArrowError::NotYetImplemented(strip("Not yet implemented: ")),
        ffi::ACU_ERR_IPC => ArrowError::IpcError(strip("Ipc error: ")),
        ffi::ACU_ERR_PARSE => ArrowError::ParseError(strip("Parser error: ")),
        ffi::ACU_ERR_PANIC_OUT_OF_BOUNDS => panic!("{full}"),
        ffi::ACU_ERR_OUT_OF_MEMORY => ArrowError::MemoryError(full),
        _ => ArrowError::ExternalError(full.into()), // ACU_ERR_CUDA / ACU_ERR_NCCL: no counterpart in the reference
    }
}
