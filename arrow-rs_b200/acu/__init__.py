"""acu — Python host-side mirror of the arrow-rs compute API over the arrow_cuda C ABI.

This is plumbing for tests and bench.py: arrays live in numpy on the host (``HostArray``,
mirroring PrimitiveArray / BooleanArray: values buffer + LSB-first validity bitmap + bit
offset + cached null_count) or in HBM (``DeviceArray``). ``Context`` exposes the reference's
function names (filter, take, add, lt, cast, sum ...) and raises ``ArrowError`` with the
reference's message text. The compute always happens in libarrow_cuda.so; nothing here
falls back to the CPU.

Reference API being mirrored: arrow/src/compute/mod.rs:20-40, arrow/src/compute/kernels.rs:20-34.
"""
import ctypes as C

import numpy as np

from . import _abi as abi
from ._abi import (ADD, ADD_WRAPPING, DISTINCT, DIV, EQ, F32, F64, GT, GT_EQ, I8, I16, I32, I64, LT, LT_EQ, MAX,
                   MIN, MUL, MUL_WRAPPING, NEQ, NOT_DISTINCT, REM, SUB, SUB_WRAPPING, SUM, U8, U16, U32, U64,
                   bitmap_bytes)

BOOL = "bool"
NP_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]


class ArrowError(Exception):
    """Mirrors arrow_schema::ArrowError; str(e) == the reference's Display text."""

    def __init__(self, status, message, index=-1, detail=None):
        super().__init__(message)
        self.status = status
        self.message = message
        self.index = index
        self.detail = detail


def pack_bits(bools, offset=0):
    """LSB-first bitmap of `bools` starting at bit `offset`; padded to whole u64 words (+8 B)."""
    bools = np.asarray(bools, dtype=bool)
    n = len(bools) + offset
    buf = np.zeros(bitmap_bytes(n) + 8, dtype=np.uint8)
    if len(bools):
        padded = np.zeros(n, dtype=bool)
        padded[offset:] = bools
        packed = np.packbits(padded, bitorder="little")
        buf[: len(packed)] = packed
    return buf


def unpack_bits(buf, offset, n):
    if n == 0:
        return np.zeros(0, dtype=bool)
    bits = np.unpackbits(np.asarray(buf, dtype=np.uint8), bitorder="little")
    return bits[offset: offset + n].astype(bool)


class HostArray:
    """Primitive or boolean Arrow array in host memory (numpy)."""

    def __init__(self, dtype, values, length, validity=None, validity_offset=0, values_offset=0, null_count=-1,
                 is_scalar=False):
        self.dtype = dtype
        self.values = values            # np array of natives, or uint8 bitmap when dtype == BOOL
        self.values_offset = values_offset
        self.validity = validity        # uint8 bitmap or None
        self.validity_offset = validity_offset
        self.length = length
        self.null_count = null_count
        self.is_scalar = is_scalar

    # -- constructors ------------------------------------------------------------------
    @staticmethod
    def from_list(dtype, items, force_validity=False, bit_offset=0, scalar=False):
        """Build from a python list; None = null (like `Int32Array::from(vec![Some(1), None])`)."""
        n = len(items)
        mask = np.array([x is not None for x in items], dtype=bool)
        if dtype == BOOL:
            vals = pack_bits([bool(x) if x is not None else False for x in items], bit_offset)
            voff = bit_offset
        else:
            npdt = NP_DTYPES[dtype]
            vals = np.zeros(n, dtype=npdt)
            for i, x in enumerate(items):
                if x is not None:
                    vals[i] = x
            voff = 0
        validity = None
        nc = 0
        if force_validity or not mask.all():
            validity = pack_bits(mask, bit_offset)
            nc = int(n - mask.sum())
        return HostArray(dtype, vals, n, validity, bit_offset if validity is not None else 0, voff, nc, scalar)

    @staticmethod
    def from_numpy(dtype, values, mask=None, bit_offset=0):
        values = np.ascontiguousarray(values, dtype=NP_DTYPES[dtype])
        validity, nc = None, 0
        if mask is not None:
            mask = np.asarray(mask, dtype=bool)
            validity = pack_bits(mask, bit_offset)
            nc = int(len(mask) - mask.sum())
        return HostArray(dtype, values, len(values), validity, bit_offset if validity is not None else 0, 0, nc)

    @staticmethod
    def bool_from_numpy(bools, mask=None, bit_offset=0, mask_offset=0):
        bools = np.asarray(bools, dtype=bool)
        validity, nc = None, 0
        if mask is not None:
            mask = np.asarray(mask, dtype=bool)
            validity = pack_bits(mask, mask_offset)
            nc = int(len(mask) - mask.sum())
        return HostArray(BOOL, pack_bits(bools, bit_offset), len(bools), validity, mask_offset, bit_offset, nc)

    def scalar(self):
        """Wrap a 1-element array as a Datum scalar (arrow-array/src/scalar.rs:128-152)."""
        assert self.length == 1
        return HostArray(self.dtype, self.values, 1, self.validity, self.validity_offset, self.values_offset,
                         self.null_count, True)

    # -- views ------------------------------------------------------------------------
    def valid_mask(self):
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        return unpack_bits(self.validity, self.validity_offset, self.length)

    def value_array(self):
        if self.dtype == BOOL:
            return unpack_bits(self.values, self.values_offset, self.length)
        return np.asarray(self.values[: self.length])

    def to_list(self):
        vals, mask = self.value_array(), self.valid_mask()
        return [(v.item() if m else None) for v, m in zip(vals, mask)]

    def slice(self, offset, length):
        """Array::slice — zero-copy: pointer/bit-offset arithmetic only."""
        if self.dtype == BOOL:
            vals, voff = self.values, self.values_offset + offset
        else:
            vals, voff = self.values[offset:], 0
        nc = -1 if self.validity is not None else 0
        return HostArray(self.dtype, vals, length, self.validity, self.validity_offset + offset if self.validity is not None else 0,
                         voff, nc, False)

    def width(self):
        return 1 if self.dtype == BOOL else abi.DTYPE_SIZE[self.dtype]


def _np_ptr(a):
    return a.ctypes.data if a is not None else None


def host_descriptor(h):
    """acu_array pointing at numpy memory (used by the oracle wrapper in tests/)."""
    d = abi.Array()
    d.values = _np_ptr(h.values)
    d.values_offset = h.values_offset
    d.validity = _np_ptr(h.validity)
    d.validity_offset = h.validity_offset
    d.len = h.length
    d.null_count = h.null_count if h.validity is not None else 0
    d.is_scalar = 1 if h.is_scalar else 0
    return d


class Utf8Column:
    """A Utf8 / Binary column on the host: offsets (np.int32 | np.int64, rows + 1), value bytes (np.uint8) and a
    HostArray carrying only the validity / length (GenericByteArray, arrow-array/src/array/byte_array.rs)."""

    def __init__(self, offsets, data, nulls):
        self.offsets, self.data, self.nulls = offsets, data, nulls

    @property
    def length(self):
        return len(self.offsets) - 1


class ViewColumn:
    """A Utf8View / BinaryView column on the host (GenericByteViewArray, arrow-array/src/array/byte_view_array.rs): `views` is an
    (n, 16) uint8 array — length u32 | 12 inline bytes, or length | 4-byte prefix | buffer index u32 | offset u32
    (arrow-data/src/byte_view.rs) — `buffers` the data buffers (uint8 arrays), `nulls` a HostArray carrying validity / length /
    scalar-ness."""

    def __init__(self, views, buffers, nulls):
        self.views, self.buffers, self.nulls = views, buffers, nulls

    @property
    def length(self):
        return self.nulls.length

    @staticmethod
    def from_values(items, block_size=64, scalar=False, garbage_under_nulls=None):
        """items: list of bytes / str / None. Long values (> 12 bytes) are appended to data buffers of `block_size` bytes
        (a new buffer is started when one is full, like GenericByteViewBuilder)."""
        n = len(items)
        views = np.zeros((n, 16), dtype=np.uint8)
        buffers, cur = [], bytearray()
        for i, it in enumerate(items):
            if it is None:
                if garbage_under_nulls is not None:
                    views[i] = garbage_under_nulls[i % len(garbage_under_nulls)]
                continue
            b = it.encode() if isinstance(it, str) else bytes(it)
            views[i, :4] = np.frombuffer(np.uint32(len(b)).tobytes(), dtype=np.uint8)
            if len(b) <= 12:
                views[i, 4:4 + len(b)] = np.frombuffer(b, dtype=np.uint8)
            else:
                if len(cur) + len(b) > block_size and len(cur):
                    buffers.append(np.frombuffer(bytes(cur), dtype=np.uint8).copy())
                    cur = bytearray()
                views[i, 4:8] = np.frombuffer(b[:4], dtype=np.uint8)
                views[i, 8:12] = np.frombuffer(np.uint32(len(buffers)).tobytes(), dtype=np.uint8)
                views[i, 12:16] = np.frombuffer(np.uint32(len(cur)).tobytes(), dtype=np.uint8)
                cur += b
        if len(cur):
            buffers.append(np.frombuffer(bytes(cur), dtype=np.uint8).copy())
        mask = np.array([it is not None for it in items], dtype=bool)
        nulls = HostArray.from_list(U8, [0 if m else None for m in mask])
        nulls.values = np.zeros(0, np.uint8)
        if scalar:
            nulls.is_scalar = True
        return ViewColumn(views, buffers, nulls)

    def values(self):
        out, mask = [], self.nulls.valid_mask()
        for i in range(self.length):
            if not mask[i]:
                out.append(None)
                continue
            ln = int(np.frombuffer(self.views[i, :4].tobytes(), dtype=np.uint32)[0])
            if ln <= 12:
                out.append(bytes(self.views[i, 4:4 + ln]))
            else:
                bi, off = (int(x) for x in np.frombuffer(self.views[i, 8:16].tobytes(), dtype=np.uint32))
                out.append(bytes(self.buffers[bi][off:off + ln]))
        return out


class DeviceArray:
    """A HostArray's buffers uploaded to HBM (DeviceBuffer pair) with the same offsets."""

    def __init__(self, ctx, dtype, length, d_values, values_offset, d_validity, validity_offset, null_count, is_scalar,
                 owned):
        self.ctx, self.dtype, self.length = ctx, dtype, length
        self.d_values, self.values_offset = d_values, values_offset
        self.d_validity, self.validity_offset = d_validity, validity_offset
        self.null_count, self.is_scalar = null_count, is_scalar
        self._owned = owned

    def descriptor(self):
        d = abi.Array()
        d.values = self.d_values
        d.values_offset = self.values_offset
        d.validity = self.d_validity
        d.validity_offset = self.validity_offset
        d.len = self.length
        d.null_count = self.null_count if self.d_validity else 0
        d.is_scalar = 1 if self.is_scalar else 0
        return d

    def free(self):
        for p in self._owned:
            self.ctx.free(p)
        self._owned = []


class Context:
    """acu_ctx wrapper: one device, one stream."""

    def __init__(self, device=0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        st = self.lib.acu_ctx_create(device, C.byref(h))
        if st != abi.OK:
            raise RuntimeError(f"acu_ctx_create(device={device}) failed with status {st}: no usable CUDA device "
                               "(arrow-cuda has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.acu_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- errors / memory -----------------------------------------------------------------
    def check(self, st):
        if st != abi.OK:
            d = self.lib.acu_last_error(self.h).contents
            raise ArrowError(st, d.message.decode(), d.index, d)

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.acu_malloc(self.h, max(int(nbytes), 1), C.byref(p)))
        return p.value

    def free(self, p):
        if p:
            self.check(self.lib.acu_free(self.h, p))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.acu_memcpy_h2d(self.h, dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, dptr, nbytes, dtype=np.uint8):
        out = np.empty(max(int(nbytes), 0) // np.dtype(dtype).itemsize, dtype=dtype)
        if out.nbytes:
            self.check(self.lib.acu_memcpy_d2h(self.h, out.ctypes.data, dptr, out.nbytes))
        return out

    def sync(self):
        self.check(self.lib.acu_ctx_sync(self.h))

    def async_begin(self):
        """Open a stream-ordered section (include/arrow_cuda.h): the supported entry points only enqueue."""
        self.check(self.lib.acu_async_begin(self.h))

    def results_fetch(self):
        """ONE D2H + ONE synchronisation; finalises the queued calls in order, raises the first error."""
        self.check(self.lib.acu_results_fetch(self.h))

    def launch_count(self):
        return self.lib.acu_launch_count(self.h)

    def upload(self, h):
        owned = []
        if h.dtype == BOOL:
            buf = np.asarray(h.values, dtype=np.uint8)
            dv = self.malloc(buf.nbytes + 8)
            self.h2d(dv, buf)
        else:
            vals = np.ascontiguousarray(h.values[: max(h.length, 1 if h.is_scalar else 0)])
            dv = self.malloc(vals.nbytes + 16)
            if vals.nbytes:
                self.h2d(dv, vals)
        owned.append(dv)
        dn = None
        if h.validity is not None:
            dn = self.malloc(h.validity.nbytes + 8)
            self.h2d(dn, h.validity)
            owned.append(dn)
        return DeviceArray(self, h.dtype, h.length, dv, h.values_offset, dn, h.validity_offset,
                           h.null_count if h.validity is not None else 0, h.is_scalar, owned)

    def alloc_out(self, nbytes_values, n_rows):
        out = abi.ArrayOut()
        out.values = self.malloc(nbytes_values + 16)
        out.validity = self.malloc(bitmap_bytes(n_rows) + 8)
        return out

    def download_out(self, out, dtype):
        n = out.len
        if dtype == BOOL:
            vals = self.d2h(out.values, bitmap_bytes(n))
        else:
            vals = self.d2h(out.values, n * abi.DTYPE_SIZE[dtype], NP_DTYPES[dtype])
        validity = self.d2h(out.validity, bitmap_bytes(n)) if out.has_validity else None
        res = HostArray(dtype, vals, n, validity, 0, 0, out.null_count if out.has_validity else 0)
        self.free(out.values)
        self.free(out.validity)
        return res

    def _free_out(self, out):
        self.free(out.values)
        self.free(out.validity)

    # -- filter (arrow-select/src/filter.rs) ----------------------------------------------
    def filter(self, values, predicate):
        """arrow::compute::filter(values, predicate) for primitive and boolean arrays."""
        dv, dp = self.upload(values), self.upload(predicate)
        plan = C.c_void_p()
        out = None
        try:
            pd = dp.descriptor()
            self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
            count = self.lib.acu_filter_plan_count(plan)
            out = self.alloc_out(count * values.width(), count)
            vd = dv.descriptor()
            if values.dtype == BOOL:
                self.check(self.lib.acu_filter_boolean(self.h, plan, C.byref(vd), C.byref(out)))
            else:
                self.check(self.lib.acu_filter_primitive(self.h, plan, values.width(), C.byref(vd), C.byref(out)))
            res, out = self.download_out(out, values.dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            dv.free()
            dp.free()

    def filter_plan(self, predicate):
        """FilterBuilder::new(predicate).optimize().build() -> (count, strategy)."""
        dp = self.upload(predicate)
        plan = C.c_void_p()
        try:
            pd = dp.descriptor()
            self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
            return self.lib.acu_filter_plan_count(plan), self.lib.acu_filter_plan_strategy(plan)
        finally:
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            dp.free()

    def filter_slices(self, predicate):
        """SlicesIterator::new(&prep_null_mask_filter(predicate)).collect() -> [(start, end)] (filter.rs:44-77)."""
        dp = self.upload(predicate)
        plan = C.c_void_p()
        out = None
        try:
            pd = dp.descriptor()
            self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
            n = C.c_int64(0)
            self.check(self.lib.acu_filter_plan_slices(self.h, plan, None, 0, C.byref(n)))
            if n.value == 0:
                return []
            out = self.malloc(n.value * 16 + 16)
            self.check(self.lib.acu_filter_plan_slices(self.h, plan, out, n.value, C.byref(n)))
            pairs = self.d2h(out, n.value * 16, np.uint64).reshape(-1, 2)
            return [(int(a), int(b)) for a, b in pairs]
        finally:
            if out:
                self.free(out)
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            dp.free()

    def chain(self, col, pred, idx, a, b, arith_op=ADD, agg_op=SUM, cmp_with=None):
        """filter(col, pred) -> take(col, idx) -> arith(a, b) -> aggregate(taken) queued in ONE stream-ordered section
        (acu_async_begin ... acu_results_fetch): one synchronisation for the five calls. With cmp_with = (op, x, y) the
        predicate is cmp(op, x, y) computed inside the section too (`pred` is ignored). Returns
        (filtered, taken, arith result, aggregate or None); raises the first error in call order at the fetch."""
        ups = [self.upload(x) for x in (col, idx, a, b)]
        dcol, didx, da, db = ups
        plan = C.c_void_p()
        outs = []
        n_pred = (cmp_with[2].length if cmp_with[1].is_scalar else cmp_with[1].length) if cmp_with else pred.length
        try:
            if cmp_with:
                cx, cy = self.upload(cmp_with[1]), self.upload(cmp_with[2])
                ups += [cx, cy]
                o_pred = self.alloc_out(bitmap_bytes(n_pred), n_pred)
                outs.append(o_pred)
            else:
                dpred = self.upload(pred)
                ups.append(dpred)
            # outputs of a filter whose plan is pending are sized for the predicate length
            o_f = self.alloc_out(max(n_pred, 1) * col.width(), n_pred)
            o_t = self.alloc_out(idx.length * col.width(), idx.length)
            n_ar = b.length if a.is_scalar and not b.is_scalar else a.length
            o_a = self.alloc_out(n_ar * a.width(), n_ar)
            outs += [o_f, o_t, o_a]
            bits, cnt = C.c_uint64(0), C.c_int64(0)
            cd, idd, ad, bd = dcol.descriptor(), didx.descriptor(), da.descriptor(), db.descriptor()
            self.async_begin()
            try:
                if cmp_with:
                    xd, yd = cx.descriptor(), cy.descriptor()
                    self.check(self.lib.acu_cmp(self.h, cmp_with[1].dtype, cmp_with[0], C.byref(xd), C.byref(yd), C.byref(o_pred)))
                    # the comparison's null count is still on the device: hand the plan a predicate without cached count is not
                    # allowed inside a section, so the fused entry point is the stream-ordered way to build a plan from a cmp
                    self.check(self.lib.acu_filter_plan_create_cmp(self.h, cmp_with[1].dtype, cmp_with[0], C.byref(xd), C.byref(yd), C.byref(plan)))
                else:
                    pd = dpred.descriptor()
                    self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
                if col.dtype == BOOL:
                    self.check(self.lib.acu_filter_boolean(self.h, plan, C.byref(cd), C.byref(o_f)))
                    self.check(self.lib.acu_take_boolean(self.h, C.byref(cd), C.byref(idd), idx.dtype, 0, C.byref(o_t)))
                else:
                    self.check(self.lib.acu_filter_primitive(self.h, plan, col.width(), C.byref(cd), C.byref(o_f)))
                    self.check(self.lib.acu_take_primitive(self.h, col.width(), C.byref(cd), C.byref(idd), idx.dtype, 0, C.byref(o_t)))
                self.check(self.lib.acu_arith(self.h, a.dtype, arith_op, C.byref(ad), C.byref(bd), C.byref(o_a)))
                has_v = (col.validity is not None and col.null_count != 0) or idx.validity is not None
                taken = abi.Array()
                taken.values, taken.values_offset = o_t.values, 0
                taken.validity, taken.validity_offset = (o_t.validity if has_v else None), 0
                taken.len, taken.null_count, taken.is_scalar = idx.length, (-1 if has_v else 0), 0
                do_agg = col.dtype != BOOL
                if do_agg:
                    self.check(self.lib.acu_aggregate(self.h, col.dtype, agg_op, C.byref(taken), C.byref(bits), C.byref(cnt)))
            except BaseException:
                try:
                    self.results_fetch()
                except ArrowError:
                    pass
                raise
            self.results_fetch()
            pred_out = self.download_out(outs.pop(0), BOOL) if cmp_with else None
            filtered = self.download_out(o_f, col.dtype)
            taken_h = self.download_out(o_t, col.dtype)
            added = self.download_out(o_a, a.dtype)
            outs = []
            agg = None
            if do_agg and cnt.value != 0:
                agg = np.array([bits.value], dtype=np.uint64).view(NP_DTYPES[col.dtype])[0].item()
            res = (filtered, taken_h, added, agg)
            return res + (pred_out,) if cmp_with else res
        finally:
            for o in outs:
                self._free_out(o)
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            for u in ups:
                u.free()

    # -- take (arrow-select/src/take.rs) ----------------------------------------------------
    def take(self, values, indices, check_bounds=False):
        dv, di = self.upload(values), self.upload(indices)
        out = self.alloc_out(indices.length * values.width(), indices.length)
        try:
            vd, idd = dv.descriptor(), di.descriptor()
            if values.dtype == BOOL:
                self.check(self.lib.acu_take_boolean(self.h, C.byref(vd), C.byref(idd), indices.dtype, int(check_bounds), C.byref(out)))
            else:
                self.check(self.lib.acu_take_primitive(self.h, values.width(), C.byref(vd), C.byref(idd), indices.dtype,
                                                       int(check_bounds), C.byref(out)))
            res, out = self.download_out(out, values.dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            dv.free()
            di.free()

    # -- variable width (Utf8) ---------------------------------------------------------------
    def take_bytes(self, offsets, data, nulls_of, indices, check_bounds=False):
        """take on a Utf8/Binary array given as (offsets np.int32/int64, data np.uint8, nulls_of HostArray
        carrying validity/len). Returns (offsets, data, nulls HostArray)."""
        ob = offsets.dtype.itemsize
        d_off = self.malloc(offsets.nbytes + 16)
        self.h2d(d_off, offsets)
        d_data = self.malloc(data.nbytes + 16)
        if data.nbytes:
            self.h2d(d_data, data)
        dn, di = self.upload(nulls_of), self.upload(indices)
        m = indices.length
        d_out_off = self.malloc((m + 1) * ob + 16)
        out = self.alloc_out(0, m)
        d_out_data = None
        try:
            total = C.c_int64(0)
            nd, idd = dn.descriptor(), di.descriptor()
            self.check(self.lib.acu_take_bytes(self.h, ob, d_off, d_data, C.byref(nd), C.byref(idd), indices.dtype,
                                               int(check_bounds), d_out_off, None, 0, C.byref(total), C.byref(out)))
            d_out_data = self.malloc(total.value + 16)
            self.check(self.lib.acu_take_bytes(self.h, ob, d_off, d_data, C.byref(nd), C.byref(idd), indices.dtype,
                                               int(check_bounds), d_out_off, d_out_data, total.value, C.byref(total), C.byref(out)))
            o = self.d2h(d_out_off, (m + 1) * ob, offsets.dtype)
            b = self.d2h(d_out_data, total.value)
            validity = self.d2h(out.validity, bitmap_bytes(m)) if out.has_validity else None
            return o, b, HostArray(U8, np.zeros(0, np.uint8), m, validity, 0, 0, out.null_count if out.has_validity else 0)
        finally:
            self._free_out(out)
            for p in (d_off, d_data, d_out_off, d_out_data):
                self.free(p)
            dn.free()
            di.free()

    def filter_bytes(self, offsets, data, nulls_of, predicate):
        ob = offsets.dtype.itemsize
        d_off = self.malloc(offsets.nbytes + 16)
        self.h2d(d_off, offsets)
        d_data = self.malloc(data.nbytes + 16)
        if data.nbytes:
            self.h2d(d_data, data)
        dn, dp = self.upload(nulls_of), self.upload(predicate)
        plan = C.c_void_p()
        d_out_off = d_out_data = None
        out = None
        try:
            pd = dp.descriptor()
            self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
            count = self.lib.acu_filter_plan_count(plan)
            d_out_off = self.malloc((count + 1) * ob + 16)
            out = self.alloc_out(0, count)
            total = C.c_int64(0)
            nd = dn.descriptor()
            self.check(self.lib.acu_filter_bytes(self.h, plan, ob, d_off, d_data, C.byref(nd), d_out_off, None, 0,
                                                 C.byref(total), C.byref(out)))
            d_out_data = self.malloc(total.value + 16)
            self.check(self.lib.acu_filter_bytes(self.h, plan, ob, d_off, d_data, C.byref(nd), d_out_off, d_out_data,
                                                 total.value, C.byref(total), C.byref(out)))
            o = self.d2h(d_out_off, (count + 1) * ob, offsets.dtype)
            b = self.d2h(d_out_data, total.value)
            validity = self.d2h(out.validity, bitmap_bytes(count)) if out.has_validity else None
            return o, b, HostArray(U8, np.zeros(0, np.uint8), count, validity, 0, 0, out.null_count if out.has_validity else 0)
        finally:
            if out is not None:
                self._free_out(out)
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            for p in (d_off, d_data, d_out_off, d_out_data):
                self.free(p)
            dn.free()
            dp.free()

    # -- RecordBatch level (filter.rs:225-244, take.rs:1123-1133) ----------------------------
    # A batch is a list of columns; a column is a HostArray (primitive / boolean) or a
    # Utf8Column(offsets, data, nulls). Results keep the column order and kinds.
    def _upload_columns(self, columns):
        cols = (abi.Column * len(columns))()
        owned = []
        for c, col in enumerate(columns):
            if isinstance(col, Utf8Column):
                d_off = self.malloc(col.offsets.nbytes + 16)
                self.h2d(d_off, col.offsets)
                d_data = self.malloc(col.data.nbytes + 16)
                if col.data.nbytes:
                    self.h2d(d_data, col.data)
                dn = self.upload(col.nulls)
                owned += [("ptr", d_off), ("ptr", d_data), ("arr", dn)]
                cols[c].kind, cols[c].width = abi.COL_BYTES, col.offsets.dtype.itemsize
                cols[c].array = dn.descriptor()
                cols[c].array.values = d_off
                cols[c].array.values_offset = 0
                cols[c].data = d_data
            else:
                dv = self.upload(col)
                owned.append(("arr", dv))
                cols[c].kind = abi.COL_BOOLEAN if col.dtype == BOOL else abi.COL_PRIMITIVE
                cols[c].width = 0 if col.dtype == BOOL else col.width()
                cols[c].array = dv.descriptor()
        return cols, owned

    def _alloc_column_outs(self, columns, rows, data_caps):
        outs = (abi.ColumnOut * len(columns))()
        for c, col in enumerate(columns):
            if isinstance(col, Utf8Column):
                outs[c].array.values = self.malloc((rows + 1) * col.offsets.dtype.itemsize + 16)
                outs[c].array.validity = self.malloc(bitmap_bytes(rows) + 8)
                outs[c].data = self.malloc(data_caps[c] + 16)
                outs[c].data_capacity = data_caps[c]
            else:
                outs[c].array.values = self.malloc((bitmap_bytes(rows) if col.dtype == BOOL else rows * col.width()) + 16)
                outs[c].array.validity = self.malloc(bitmap_bytes(rows) + 8)
        return outs

    def _download_columns(self, columns, outs):
        res = []
        for c, col in enumerate(columns):
            o = outs[c]
            n = o.array.len
            validity = self.d2h(o.array.validity, bitmap_bytes(n)) if o.array.has_validity else None
            nc = o.array.null_count if o.array.has_validity else 0
            if isinstance(col, Utf8Column):
                offs = self.d2h(o.array.values, (n + 1) * col.offsets.dtype.itemsize, col.offsets.dtype)
                data = self.d2h(o.data, o.data_len)
                res.append(Utf8Column(offs, data, HostArray(U8, np.zeros(0, np.uint8), n, validity, 0, 0, nc)))
            elif col.dtype == BOOL:
                res.append(HostArray(BOOL, self.d2h(o.array.values, bitmap_bytes(n)), n, validity, 0, 0, nc))
            else:
                res.append(HostArray(col.dtype, self.d2h(o.array.values, n * col.width(), NP_DTYPES[col.dtype]), n, validity, 0, 0, nc))
        return res

    def _free_columns(self, owned, outs):
        for kind, x in owned:
            if kind == "ptr":
                self.free(x)
            else:
                x.free()
        if outs is not None:
            for o in outs:
                for p in (o.array.values, o.array.validity, o.data):
                    if p:
                        self.free(p)

    # -- Arrow IPC stream -> HBM (arrow-ipc/src/reader.rs StreamReader) ---------------------------
    def ipc_read_stream(self, stream_bytes, on_batch=None):
        """StreamReader over an in-memory IPC stream: every RecordBatch body is ONE host->device copy and its columns are
        views into that device buffer. Returns (schema, batches): schema = [(name, kind, width, dtype, nullable)], batches =
        lists of downloaded columns (HostArray / Utf8Column) — or whatever `on_batch(columns_descriptor_array, rows)` returns
        when given (the descriptors are only valid inside the callback)."""
        buf = np.frombuffer(bytes(stream_bytes), dtype=np.uint8).copy()
        h = C.c_void_p()
        n = C.c_int32(0)
        self.check(self.lib.acu_ipc_stream_open(self.h, buf.ctypes.data, len(buf), C.byref(h), C.byref(n)))
        try:
            schema = []
            for i in range(n.value):
                kind, width, dtype, nullable, name = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_char_p()
                self.lib.acu_ipc_stream_field(h, i, C.byref(kind), C.byref(width), C.byref(dtype), C.byref(nullable), C.byref(name))
                schema.append((name.value.decode(), kind.value, width.value, dtype.value, bool(nullable.value)))
            batches = []
            cols = (abi.Column * max(n.value, 1))()
            while True:
                rows = C.c_int64(0)
                self.check(self.lib.acu_ipc_stream_next(self.h, h, cols, C.byref(rows)))
                if rows.value < 0:
                    break
                if on_batch is not None:
                    batches.append(on_batch(cols, rows.value))
                    continue
                out = []
                for i, (name, kind, width, dtype, _) in enumerate(schema):
                    c, m = cols[i], rows.value
                    validity = self.d2h(c.array.validity, bitmap_bytes(m)) if c.array.validity else None
                    nc = c.array.null_count if c.array.validity else 0
                    if kind == abi.COL_BYTES:
                        odt = np.int32 if width == 4 else np.int64
                        offs = self.d2h(c.array.values, (m + 1) * width, odt)
                        data = self.d2h(c.data, int(offs[-1]) if m else 0)
                        out.append(Utf8Column(offs, data, HostArray(U8, np.zeros(0, np.uint8), m, validity, 0, 0, nc)))
                    elif kind == abi.COL_BOOLEAN:
                        out.append(HostArray(BOOL, self.d2h(c.array.values, bitmap_bytes(m)), m, validity, 0, 0, nc))
                    else:
                        out.append(HostArray(dtype, self.d2h(c.array.values, m * width, NP_DTYPES[dtype]), m, validity, 0, 0, nc))
                batches.append(out)
            return schema, batches
        finally:
            self.lib.acu_ipc_stream_close(self.h, h)

    # -- concat / concat_batches (arrow-select/src/concat.rs:495-640) ---------------------------
    def concat(self, columns):
        """arrow::compute::concat(&[..]): columns = HostArrays or Utf8Columns of one type."""
        cols, owned = self._upload_columns(columns) if columns else ((abi.Column * 1)(), [])
        outs = None
        try:
            rows = sum(c.length for c in columns)
            proto = columns[:1] if columns else [HostArray(U8, np.zeros(0, np.uint8), 0)]
            caps = [sum(int(c.data.nbytes) for c in columns)] if columns and isinstance(columns[0], Utf8Column) else [0]
            outs = self._alloc_column_outs(proto, max(rows, 1), caps)
            self.check(self.lib.acu_concat(self.h, len(columns), cols, outs))
            return self._download_columns(proto, outs)[0]
        finally:
            self._free_columns(owned, outs)

    def concat_batches(self, batches):
        """arrow::compute::concat_batches(schema, batches): batches = lists of columns (same schema)."""
        ncols = len(batches[0]) if batches else 0
        flat = [c for b in batches for c in b]
        cols, owned = self._upload_columns(flat) if flat else ((abi.Column * 1)(), [])
        outs = None
        try:
            rows = sum(b[0].length for b in batches) if ncols else 0
            proto = list(batches[0]) if batches else []
            caps = [sum(int(b[c].data.nbytes) for b in batches) if isinstance(proto[c], Utf8Column) else 0 for c in range(ncols)]
            outs = self._alloc_column_outs(proto, max(rows, 1), caps) if ncols else (abi.ColumnOut * 1)()
            n = C.c_int64(0)
            self.check(self.lib.acu_concat_batches(self.h, len(batches), ncols, cols, outs, C.byref(n)))
            return self._download_columns(proto, outs) if ncols else []
        finally:
            self._free_columns(owned, outs if ncols else None)

    def filter_record_batch(self, columns, predicate):
        """arrow::compute::filter_record_batch: one plan, every column, one synchronisation."""
        cols, owned = self._upload_columns(columns)
        dp = self.upload(predicate)
        plan = C.c_void_p()
        outs = None
        try:
            pd = dp.descriptor()
            self.check(self.lib.acu_filter_plan_create(self.h, C.byref(pd), C.byref(plan)))
            count = self.lib.acu_filter_plan_count(plan)
            caps = [int(col.data.nbytes) if isinstance(col, Utf8Column) else 0 for col in columns]
            outs = self._alloc_column_outs(columns, count, caps)
            self.check(self.lib.acu_filter_record_batch(self.h, plan, len(columns), cols, outs))
            return self._download_columns(columns, outs)
        finally:
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            self._free_columns(owned, outs)
            dp.free()

    def take_record_batch(self, columns, indices, check_bounds=False, data_capacity=None):
        """arrow::compute::take_record_batch / take_arrays."""
        cols, owned = self._upload_columns(columns)
        di = self.upload(indices)
        outs = None
        try:
            m = indices.length
            caps = []
            for col in columns:
                if isinstance(col, Utf8Column):
                    lens = np.diff(col.offsets.astype(np.int64)) if len(col.offsets) > 1 else np.zeros(0, np.int64)
                    caps.append(int(data_capacity) if data_capacity is not None else int((lens.max() if lens.size else 0) * m))
                else:
                    caps.append(0)
            outs = self._alloc_column_outs(columns, m, caps)
            idd = di.descriptor()
            self.check(self.lib.acu_take_record_batch(self.h, len(columns), cols, C.byref(idd), indices.dtype, int(check_bounds), outs))
            return self._download_columns(columns, outs)
        finally:
            self._free_columns(owned, outs)
            di.free()

    def aggregate_columns(self, ops, columns):
        """[sum|min|max](column) for several primitive columns with one synchronisation -> [(value|None)]."""
        n = len(columns)
        das = [self.upload(c) for c in columns]
        try:
            arrs = (abi.Array * n)(*[d.descriptor() for d in das])
            dts = (C.c_int32 * n)(*[c.dtype for c in columns])
            opv = (C.c_int32 * n)(*ops)
            bits, cnts = (C.c_uint64 * n)(), (C.c_int64 * n)()
            self.check(self.lib.acu_aggregate_columns(self.h, n, dts, opv, arrs, bits, cnts))
            out = []
            for i, c in enumerate(columns):
                if cnts[i] == 0:
                    out.append(None)
                else:
                    raw = np.array([bits[i]], dtype=np.uint64).view(np.uint8)[: abi.DTYPE_SIZE[c.dtype]]
                    out.append(raw.view(NP_DTYPES[c.dtype])[0].item())
            return out
        finally:
            for d in das:
                d.free()

    # -- numeric (arrow-arith/src/numeric.rs) -----------------------------------------------
    def arith(self, op, a, b):
        assert a.dtype == b.dtype
        n = b.length if a.is_scalar and not b.is_scalar else a.length
        da, db = self.upload(a), self.upload(b)
        out = self.alloc_out(n * a.width(), n)
        try:
            ad, bd = da.descriptor(), db.descriptor()
            self.check(self.lib.acu_arith(self.h, a.dtype, op, C.byref(ad), C.byref(bd), C.byref(out)))
            res, out = self.download_out(out, a.dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            da.free()
            db.free()

    def add(self, a, b): return self.arith(ADD, a, b)
    def add_wrapping(self, a, b): return self.arith(ADD_WRAPPING, a, b)
    def sub(self, a, b): return self.arith(SUB, a, b)
    def sub_wrapping(self, a, b): return self.arith(SUB_WRAPPING, a, b)
    def mul(self, a, b): return self.arith(MUL, a, b)
    def mul_wrapping(self, a, b): return self.arith(MUL_WRAPPING, a, b)
    def div(self, a, b): return self.arith(DIV, a, b)
    def rem(self, a, b): return self.arith(REM, a, b)

    def neg(self, a, checked=True):
        da = self.upload(a)
        out = self.alloc_out(a.length * a.width(), a.length)
        try:
            ad = da.descriptor()
            self.check(self.lib.acu_neg(self.h, a.dtype, int(checked), C.byref(ad), C.byref(out)))
            res, out = self.download_out(out, a.dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            da.free()

    def neg_wrapping(self, a): return self.neg(a, checked=False)

    # -- cmp (arrow-ord/src/cmp.rs) -----------------------------------------------------------
    def cmp(self, op, a, b):
        assert a.dtype == b.dtype
        n = b.length if a.is_scalar else a.length
        da, db = self.upload(a), self.upload(b)
        out = self.alloc_out(bitmap_bytes(n), n)
        try:
            ad, bd = da.descriptor(), db.descriptor()
            self.check(self.lib.acu_cmp(self.h, a.dtype, op, C.byref(ad), C.byref(bd), C.byref(out)))
            res, out = self.download_out(out, BOOL), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            da.free()
            db.free()

    # -- cmp on Utf8 / Binary and Utf8View / BinaryView operands (cmp.rs:783-898) -----------------
    def _upload_nulls(self, nulls, owned):
        d = abi.Array()
        d.len, d.is_scalar = nulls.length, 1 if nulls.is_scalar else 0
        d.validity_offset = nulls.validity_offset
        d.null_count = nulls.null_count if nulls.validity is not None else 0
        if nulls.validity is not None:
            dn = self.malloc(nulls.validity.nbytes + 8)
            self.h2d(dn, nulls.validity)
            owned.append(dn)
            d.validity = dn
        return d

    def cmp_bytes(self, op, a, b):
        """a, b: Utf8Column (offsets, data, nulls); nulls.is_scalar marks a Datum scalar."""
        assert a.offsets.dtype == b.offsets.dtype
        owned = []
        n = max(a.nulls.length if not a.nulls.is_scalar else 0, b.nulls.length if not b.nulls.is_scalar else 0, 1)
        out = self.alloc_out(bitmap_bytes(n), n)
        try:
            descs = []
            for col in (a, b):
                d = abi.BytesArray()
                d_off, d_data = self.malloc(col.offsets.nbytes + 16), self.malloc(col.data.nbytes + 16)
                owned += [d_off, d_data]
                self.h2d(d_off, col.offsets)
                if col.data.nbytes:
                    self.h2d(d_data, col.data)
                d.offsets, d.data, d.nulls = d_off, d_data, self._upload_nulls(col.nulls, owned)
                descs.append(d)
            self.check(self.lib.acu_cmp_bytes(self.h, a.offsets.dtype.itemsize, op, C.byref(descs[0]), C.byref(descs[1]), C.byref(out)))
            res, out = self.download_out(out, BOOL), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            for p in owned:
                self.free(p)

    def cmp_view(self, op, a, b):
        """a, b: ViewColumn."""
        owned = []
        n = max(a.length if not a.nulls.is_scalar else 0, b.length if not b.nulls.is_scalar else 0, 1)
        out = self.alloc_out(bitmap_bytes(n), n)
        try:
            descs, keep = [], []
            for col in (a, b):
                d = abi.ViewArray()
                views = np.ascontiguousarray(col.views)
                d_views = self.malloc(views.nbytes + 16)
                owned.append(d_views)
                if views.nbytes:
                    self.h2d(d_views, views)
                ptrs = []
                for buf in col.buffers:
                    db = self.malloc(buf.nbytes + 16)
                    owned.append(db)
                    self.h2d(db, buf)
                    ptrs.append(db)
                table = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
                keep.append(table)
                d.views, d.buffers, d.n_buffers = d_views, table, len(ptrs)
                d.nulls = self._upload_nulls(col.nulls, owned)
                descs.append(d)
            self.check(self.lib.acu_cmp_byte_view(self.h, op, C.byref(descs[0]), C.byref(descs[1]), C.byref(out)))
            res, out = self.download_out(out, BOOL), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            for p in owned:
                self.free(p)

    # -- fused compare -> filter (cmp.rs:220-382 feeding filter.rs:254-273) -------------------
    def filter_cmp(self, values, op, a, b):
        """filter(values, &cmp::op(a, b)?) with the predicate never materialised: the comparison writes the filter plan."""
        assert a.dtype == b.dtype
        dv, da, db = self.upload(values), self.upload(a), self.upload(b)
        plan = C.c_void_p()
        out = None
        try:
            ad, bd = da.descriptor(), db.descriptor()
            self.check(self.lib.acu_filter_plan_create_cmp(self.h, a.dtype, op, C.byref(ad), C.byref(bd), C.byref(plan)))
            count = self.lib.acu_filter_plan_count(plan)
            out = self.alloc_out(count * values.width(), count)
            vd = dv.descriptor()
            if values.dtype == BOOL:
                self.check(self.lib.acu_filter_boolean(self.h, plan, C.byref(vd), C.byref(out)))
            else:
                self.check(self.lib.acu_filter_primitive(self.h, plan, values.width(), C.byref(vd), C.byref(out)))
            strategy = self.lib.acu_filter_plan_strategy(plan)
            res, out = self.download_out(out, values.dtype), None
            return res, (count, strategy)
        finally:
            if out is not None:
                self._free_out(out)
            if plan:
                self.lib.acu_filter_plan_destroy(self.h, plan)
            dv.free()
            da.free()
            db.free()

    # -- nullif / zip (arrow-select/src/nullif.rs, zip.rs) ------------------------------------
    def nullif(self, left, right):
        """arrow::compute::nullif(left, right): same values, validity &= !(right is Some(true))."""
        dl, dr = self.upload(left), self.upload(right)
        out = self.alloc_out(0, max(left.length, 1))
        try:
            ld, rd = dl.descriptor(), dr.descriptor()
            self.check(self.lib.acu_nullif(self.h, C.byref(ld), C.byref(rd), C.byref(out)))
            n = out.len
            validity = self.d2h(out.validity, bitmap_bytes(n)) if out.has_validity else None
            if n == 0:  # the array is returned as it is
                return left
            # the result shares left's value buffer (logical slice starting at row 0)
            vals = left.values if left.dtype == BOOL else left.values[:n]
            return HostArray(left.dtype, vals, n, validity, 0, left.values_offset if left.dtype == BOOL else 0,
                             out.null_count if out.has_validity else 0)
        finally:
            self._free_out(out)
            dl.free()
            dr.free()

    def zip(self, mask, truthy, falsy):
        """arrow::compute::zip(mask, truthy, falsy) for primitive arrays / scalars."""
        assert truthy.dtype == falsy.dtype and truthy.dtype != BOOL
        dm, dt, df = self.upload(mask), self.upload(truthy), self.upload(falsy)
        n = mask.length
        out = self.alloc_out(n * truthy.width(), max(n, 1))
        try:
            md, td, fd = dm.descriptor(), dt.descriptor(), df.descriptor()
            self.check(self.lib.acu_zip(self.h, truthy.width(), C.byref(md), C.byref(td), C.byref(fd), C.byref(out)))
            res, out = self.download_out(out, truthy.dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            dm.free()
            dt.free()
            df.free()

    # -- boolean (arrow-arith/src/boolean.rs) -------------------------------------------------
    def boolean(self, op, a, b=None):
        da = self.upload(a)
        db = self.upload(b) if b is not None else None
        n = max(a.length, 1)
        out = self.alloc_out(bitmap_bytes(n), n)
        try:
            ad = da.descriptor()
            bd = db.descriptor() if db is not None else None
            self.check(self.lib.acu_boolean(self.h, op, C.byref(ad), C.byref(bd) if bd is not None else None, C.byref(out)))
            res, out = self.download_out(out, BOOL), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            da.free()
            if db is not None:
                db.free()

    def and_(self, a, b): return self.boolean(abi.BOOL_AND, a, b)
    def or_(self, a, b): return self.boolean(abi.BOOL_OR, a, b)
    def and_not(self, a, b): return self.boolean(abi.BOOL_AND_NOT, a, b)
    def and_kleene(self, a, b): return self.boolean(abi.BOOL_AND_KLEENE, a, b)
    def or_kleene(self, a, b): return self.boolean(abi.BOOL_OR_KLEENE, a, b)
    def not_(self, a): return self.boolean(abi.BOOL_NOT, a)
    def is_null(self, a): return self.boolean(abi.BOOL_IS_NULL, a)
    def is_not_null(self, a): return self.boolean(abi.BOOL_IS_NOT_NULL, a)

    def eq(self, a, b): return self.cmp(EQ, a, b)
    def neq(self, a, b): return self.cmp(NEQ, a, b)
    def lt(self, a, b): return self.cmp(LT, a, b)
    def lt_eq(self, a, b): return self.cmp(LT_EQ, a, b)
    def gt(self, a, b): return self.cmp(GT, a, b)
    def gt_eq(self, a, b): return self.cmp(GT_EQ, a, b)
    def distinct(self, a, b): return self.cmp(DISTINCT, a, b)
    def not_distinct(self, a, b): return self.cmp(NOT_DISTINCT, a, b)

    # -- cast (arrow-cast/src/cast/mod.rs) -----------------------------------------------------
    def cast(self, a, to_dtype, safe=True):
        da = self.upload(a)
        out = self.alloc_out(a.length * abi.DTYPE_SIZE[to_dtype], a.length)
        try:
            ad = da.descriptor()
            self.check(self.lib.acu_cast_numeric(self.h, a.dtype, to_dtype, int(safe), C.byref(ad), C.byref(out)))
            res, out = self.download_out(out, to_dtype), None
            return res
        finally:
            if out is not None:
                self._free_out(out)
            da.free()

    # -- aggregate (arrow-arith/src/aggregate.rs) ----------------------------------------------
    def aggregate(self, op, a):
        da = self.upload(a)
        try:
            bits, cnt = C.c_uint64(0), C.c_int64(0)
            ad = da.descriptor()
            self.check(self.lib.acu_aggregate(self.h, a.dtype, op, C.byref(ad), C.byref(bits), C.byref(cnt)))
            if cnt.value == 0:
                return None
            return np.array([bits.value], dtype=np.uint64).view(NP_DTYPES[a.dtype])[0].item() if abi.DTYPE_SIZE[a.dtype] == 8 \
                else np.array([bits.value], dtype=np.uint64).view(np.uint8)[: abi.DTYPE_SIZE[a.dtype]].view(NP_DTYPES[a.dtype])[0].item()
        finally:
            da.free()

    def sum_checked(self, a):
        """arrow::compute::sum_checked (aggregate.rs:897): the in-order checked fold; raises ArrowError on overflow."""
        da = self.upload(a)
        try:
            bits, cnt = C.c_uint64(0), C.c_int64(0)
            ad = da.descriptor()
            self.check(self.lib.acu_sum_checked(self.h, a.dtype, C.byref(ad), C.byref(bits), C.byref(cnt)))
            if cnt.value == 0:
                return None
            raw = np.array([bits.value], dtype=np.uint64).view(np.uint8)[: abi.DTYPE_SIZE[a.dtype]]
            return raw.view(NP_DTYPES[a.dtype])[0].item()
        finally:
            da.free()

    def sum(self, a): return self.aggregate(SUM, a)
    def min(self, a): return self.aggregate(MIN, a)
    def max(self, a): return self.aggregate(MAX, a)
