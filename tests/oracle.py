"""ctypes wrapper of oracle/liboracle.so with the same method names as acu.Context.

TEST INFRASTRUCTURE: the oracle is the CPU restatement of the reference (see oracle/oracle.cpp);
it is the checker, never the thing measured or shipped.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import acu
from acu import _abi as abi
from acu import BOOL, HostArray, ArrowError, bitmap_bytes

ORACLE_DIR = os.path.join(abi.REPO, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

P = C.POINTER
vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_LIB):
            build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        lib.orc_last_error.restype = P(abi.ErrorDetail)
        lib.orc_last_error.argtypes = []
        sigs = {
            "orc_bitmap_count": [vp, i64, vp, i64, i64, P(i64)],
            "orc_filter_primitive": [P(abi.Array), i32, P(abi.Array), P(abi.ArrayOut), P(i64), P(i32)],
            "orc_filter_boolean": [P(abi.Array), P(abi.Array), P(abi.ArrayOut), P(i64)],
            "orc_filter_bytes": [P(abi.Array), i32, vp, vp, P(abi.Array), vp, vp, i64, P(i64), P(abi.ArrayOut), P(i64)],
            "orc_take_primitive": [i32, P(abi.Array), P(abi.Array), i32, i32, P(abi.ArrayOut)],
            "orc_take_boolean": [P(abi.Array), P(abi.Array), i32, i32, P(abi.ArrayOut)],
            "orc_take_bytes": [i32, vp, vp, P(abi.Array), P(abi.Array), i32, i32, vp, vp, i64, P(i64), P(abi.ArrayOut)],
            "orc_arith": [i32, i32, P(abi.Array), P(abi.Array), P(abi.ArrayOut)],
            "orc_neg": [i32, i32, P(abi.Array), P(abi.ArrayOut)],
            "orc_cmp": [i32, i32, P(abi.Array), P(abi.Array), P(abi.ArrayOut)],
            "orc_boolean": [i32, P(abi.Array), P(abi.Array), P(abi.ArrayOut)],
            "orc_cast_numeric": [i32, i32, i32, P(abi.Array), P(abi.ArrayOut)],
            "orc_aggregate": [i32, i32, P(abi.Array), i32, P(u64), P(i64)],
            "orc_sum_checked": [i32, P(abi.Array), P(u64), P(i64)],
            "orc_cmp_bytes": [i32, i32, vp, vp, P(abi.Array), vp, vp, P(abi.Array), P(abi.ArrayOut)],
            "orc_cmp_byte_view": [i32, vp, P(vp), i32, P(abi.Array), vp, P(vp), i32, P(abi.Array), P(abi.ArrayOut)],
            "orc_concat": [i32, P(abi.Column), P(abi.ColumnOut)],
            "orc_nullif": [P(abi.Array), P(abi.Array), P(abi.ArrayOut)],
            "orc_view_fit": [vp, i64, i64, P(i64), P(i64)],
            "orc_view_rebase": [vp, i64, C.c_uint32, vp],
            "orc_zip": [i32, P(abi.Array), P(abi.Array), P(abi.Array), P(abi.ArrayOut)],
            "orc_generate_values": [i32, u64, i64, u64, vp, i64],
            "orc_generate_bits": [u64, i64, C.c_double, vp, i64],
            "orc_bench_create": [i64, i64, i32, P(u64), C.c_double, C.c_double, i32, P(vp), P(C.c_double)],
            "orc_bench_step": [vp, P(C.c_double), P(u64), P(i64)],
            "orc_bench_check": [vp, P(u64)],
        }
        lib.orc_bench_threads.restype = i32
        lib.orc_bench_threads.argtypes = [vp]
        lib.orc_bench_destroy.restype = None
        lib.orc_bench_destroy.argtypes = [vp]
        for name, args in sigs.items():
            fn = getattr(lib, name)
            fn.restype = i32
            fn.argtypes = args
        self.lib = lib

    # ------------------------------------------------------------------------------------
    def check(self, st):
        if st != abi.OK:
            d = self.lib.orc_last_error().contents
            raise ArrowError(st, d.message.decode(), d.index, d)

    @staticmethod
    def _out(nbytes_values, n_rows):
        out = abi.ArrayOut()
        vals = np.zeros(nbytes_values + 16, dtype=np.uint8)
        valid = np.zeros(bitmap_bytes(n_rows) + 8, dtype=np.uint8)
        out.values = vals.ctypes.data
        out.validity = valid.ctypes.data
        return out, vals, valid

    @staticmethod
    def _result(out, vals, valid, dtype):
        n = out.len
        if dtype == BOOL:
            v = vals[: bitmap_bytes(n)].copy()
        else:
            v = vals[: n * abi.DTYPE_SIZE[dtype]].copy().view(acu.NP_DTYPES[dtype])
        validity = valid[: bitmap_bytes(n)].copy() if out.has_validity else None
        return HostArray(dtype, v, n, validity, 0, 0, out.null_count if out.has_validity else 0)

    # -- filter ----------------------------------------------------------------------------
    def filter(self, values, predicate):
        cap = max(values.length, predicate.length)
        out, vals, valid = self._out(cap * values.width(), cap)
        pd, vd = acu.host_descriptor(predicate), acu.host_descriptor(values)
        cnt = i64(0)
        if values.dtype == BOOL:
            self.check(self.lib.orc_filter_boolean(C.byref(pd), C.byref(vd), C.byref(out), C.byref(cnt)))
        else:
            strat = i32(0)
            self.check(self.lib.orc_filter_primitive(C.byref(pd), values.width(), C.byref(vd), C.byref(out), C.byref(cnt), C.byref(strat)))
        return self._result(out, vals, valid, values.dtype)

    def filter_slices(self, predicate):
        """SlicesIterator::new(&prep_null_mask_filter(predicate)).collect() -> [(start, end)]."""
        self.lib.orc_filter_slices.restype = i64
        self.lib.orc_filter_slices.argtypes = [P(abi.Array), vp, i64]
        pd = acu.host_descriptor(predicate)
        cap = predicate.length // 2 + 2
        out = np.zeros(2 * cap, dtype=np.uint64)
        n = self.lib.orc_filter_slices(C.byref(pd), out.ctypes.data, cap)
        return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n)]

    def filter_plan(self, predicate):
        out, vals, valid = self._out(predicate.length, predicate.length)
        dummy = HostArray(acu.U8, np.zeros(predicate.length + 1, np.uint8), predicate.length)
        pd, vd = acu.host_descriptor(predicate), acu.host_descriptor(dummy)
        cnt, strat = i64(0), i32(0)
        self.check(self.lib.orc_filter_primitive(C.byref(pd), 1, C.byref(vd), C.byref(out), C.byref(cnt), C.byref(strat)))
        return cnt.value, strat.value

    def filter_bytes(self, offsets, data, nulls_of, predicate):
        ob = offsets.dtype.itemsize
        n = predicate.length
        out, vals, valid = self._out(0, n)
        out_off = np.zeros(n + 2, dtype=offsets.dtype)
        out_data = np.zeros(len(data) + 16, dtype=np.uint8)
        pd, nd = acu.host_descriptor(predicate), acu.host_descriptor(nulls_of)
        total, cnt = i64(0), i64(0)
        self.check(self.lib.orc_filter_bytes(C.byref(pd), ob, offsets.ctypes.data, data.ctypes.data, C.byref(nd),
                                             out_off.ctypes.data, out_data.ctypes.data, len(data), C.byref(total),
                                             C.byref(out), C.byref(cnt)))
        m = out.len
        validity = valid[: bitmap_bytes(m)].copy() if out.has_validity else None
        return out_off[: m + 1].copy(), out_data[: total.value].copy(), HostArray(acu.U8, np.zeros(0, np.uint8), m, validity, 0, 0,
                                                                                  out.null_count if out.has_validity else 0)

    # -- take ------------------------------------------------------------------------------
    def take(self, values, indices, check_bounds=False):
        out, vals, valid = self._out(indices.length * values.width(), indices.length)
        vd, idd = acu.host_descriptor(values), acu.host_descriptor(indices)
        if values.dtype == BOOL:
            self.check(self.lib.orc_take_boolean(C.byref(vd), C.byref(idd), indices.dtype, int(check_bounds), C.byref(out)))
        else:
            self.check(self.lib.orc_take_primitive(values.width(), C.byref(vd), C.byref(idd), indices.dtype, int(check_bounds), C.byref(out)))
        return self._result(out, vals, valid, values.dtype)

    def take_bytes(self, offsets, data, nulls_of, indices, check_bounds=False):
        ob = offsets.dtype.itemsize
        m = indices.length
        out, vals, valid = self._out(0, m)
        out_off = np.zeros(m + 2, dtype=offsets.dtype)
        nd, idd = acu.host_descriptor(nulls_of), acu.host_descriptor(indices)
        total = i64(0)
        self.check(self.lib.orc_take_bytes(ob, offsets.ctypes.data, data.ctypes.data, C.byref(nd), C.byref(idd), indices.dtype,
                                           int(check_bounds), out_off.ctypes.data, None, 0, C.byref(total), C.byref(out)))
        out_data = np.zeros(total.value + 16, dtype=np.uint8)
        self.check(self.lib.orc_take_bytes(ob, offsets.ctypes.data, data.ctypes.data, C.byref(nd), C.byref(idd), indices.dtype,
                                           int(check_bounds), out_off.ctypes.data, out_data.ctypes.data, total.value, C.byref(total), C.byref(out)))
        validity = valid[: bitmap_bytes(m)].copy() if out.has_validity else None
        return out_off[: m + 1].copy(), out_data[: total.value].copy(), HostArray(acu.U8, np.zeros(0, np.uint8), m, validity, 0, 0,
                                                                                  out.null_count if out.has_validity else 0)

    # -- numeric -----------------------------------------------------------------------------
    def arith(self, op, a, b):
        n = b.length if a.is_scalar and not b.is_scalar else a.length
        n = max(n, a.length if not a.is_scalar else 0, b.length if not b.is_scalar else 0, 1)
        out, vals, valid = self._out(n * a.width(), n)
        ad, bd = acu.host_descriptor(a), acu.host_descriptor(b)
        self.check(self.lib.orc_arith(a.dtype, op, C.byref(ad), C.byref(bd), C.byref(out)))
        return self._result(out, vals, valid, a.dtype)

    def add(self, a, b): return self.arith(abi.ADD, a, b)
    def add_wrapping(self, a, b): return self.arith(abi.ADD_WRAPPING, a, b)
    def sub(self, a, b): return self.arith(abi.SUB, a, b)
    def sub_wrapping(self, a, b): return self.arith(abi.SUB_WRAPPING, a, b)
    def mul(self, a, b): return self.arith(abi.MUL, a, b)
    def mul_wrapping(self, a, b): return self.arith(abi.MUL_WRAPPING, a, b)
    def div(self, a, b): return self.arith(abi.DIV, a, b)
    def rem(self, a, b): return self.arith(abi.REM, a, b)

    def neg(self, a, checked=True):
        out, vals, valid = self._out(a.length * a.width(), a.length)
        ad = acu.host_descriptor(a)
        self.check(self.lib.orc_neg(a.dtype, int(checked), C.byref(ad), C.byref(out)))
        return self._result(out, vals, valid, a.dtype)

    def neg_wrapping(self, a): return self.neg(a, checked=False)

    # -- cmp ---------------------------------------------------------------------------------
    def cmp(self, op, a, b):
        n = max(a.length, b.length, 1)
        out, vals, valid = self._out(bitmap_bytes(n), n)
        ad, bd = acu.host_descriptor(a), acu.host_descriptor(b)
        self.check(self.lib.orc_cmp(a.dtype, op, C.byref(ad), C.byref(bd), C.byref(out)))
        return self._result(out, vals, valid, BOOL)

    # -- concat ---------------------------------------------------------------------------------
    def concat(self, columns):
        n = len(columns)
        cols = (abi.Column * max(n, 1))()
        keep = []
        for c, col in enumerate(columns):
            if isinstance(col, acu.Utf8Column):
                cols[c].kind, cols[c].width = abi.COL_BYTES, col.offsets.dtype.itemsize
                cols[c].array = acu.host_descriptor(col.nulls)
                cols[c].array.values = col.offsets.ctypes.data
                cols[c].data = col.data.ctypes.data
            else:
                cols[c].kind = abi.COL_BOOLEAN if col.dtype == BOOL else abi.COL_PRIMITIVE
                cols[c].width = 0 if col.dtype == BOOL else col.width()
                cols[c].array = acu.host_descriptor(col)
            keep.append(col)
        rows = sum(col.length for col in columns)
        out = abi.ColumnOut()
        proto = columns[0] if columns else HostArray(acu.U8, np.zeros(0, np.uint8), 0)
        valid = np.zeros(bitmap_bytes(max(rows, 1)) + 8, dtype=np.uint8)
        out.array.validity = valid.ctypes.data
        if isinstance(proto, acu.Utf8Column):
            offs = np.zeros(rows + 2, dtype=proto.offsets.dtype)
            cap = sum(int(c.data.nbytes) for c in columns)
            data = np.zeros(cap + 16, dtype=np.uint8)
            out.array.values, out.data, out.data_capacity = offs.ctypes.data, data.ctypes.data, cap
        else:
            vals = np.zeros((bitmap_bytes(max(rows, 1)) if proto.dtype == BOOL else rows * proto.width()) + 16, dtype=np.uint8)
            out.array.values = vals.ctypes.data
        self.check(self.lib.orc_concat(n, cols, C.byref(out)))
        m = out.array.len
        validity = valid[: bitmap_bytes(m)].copy() if out.array.has_validity else None
        nc = out.array.null_count if out.array.has_validity else 0
        if isinstance(proto, acu.Utf8Column):
            return acu.Utf8Column(offs[: m + 1].copy(), data[: out.data_len].copy(), HostArray(acu.U8, np.zeros(0, np.uint8), m, validity, 0, 0, nc))
        if proto.dtype == BOOL:
            return HostArray(BOOL, vals[: bitmap_bytes(m)].copy(), m, validity, 0, 0, nc)
        return HostArray(proto.dtype, vals[: m * proto.width()].copy().view(acu.NP_DTYPES[proto.dtype]), m, validity, 0, 0, nc)

    def concat_batches(self, batches):
        ncols = len(batches[0]) if batches else 0
        return [self.concat([b[c] for b in batches]) for c in range(ncols)]

    # -- cmp on Utf8 / Binary (Utf8Column) and Utf8View / BinaryView (ViewColumn) operands -----------
    def cmp_bytes(self, op, a, b):
        """a, b: acu.Utf8Column (offsets, data, nulls); nulls.is_scalar marks a Datum scalar."""
        assert a.offsets.dtype == b.offsets.dtype
        n = max(a.nulls.length if not a.nulls.is_scalar else 0, b.nulls.length if not b.nulls.is_scalar else 0, 1)
        out, vals, valid = self._out(bitmap_bytes(n), n)
        ad, bd = acu.host_descriptor(a.nulls), acu.host_descriptor(b.nulls)
        self.check(self.lib.orc_cmp_bytes(a.offsets.dtype.itemsize, op, a.offsets.ctypes.data, a.data.ctypes.data, C.byref(ad),
                                          b.offsets.ctypes.data, b.data.ctypes.data, C.byref(bd), C.byref(out)))
        return self._result(out, vals, valid, BOOL)

    def cmp_view(self, op, a, b):
        """a, b: acu.ViewColumn."""
        n = max(a.length if not a.nulls.is_scalar else 0, b.length if not b.nulls.is_scalar else 0, 1)
        out, vals, valid = self._out(bitmap_bytes(n), n)
        ad, bd = acu.host_descriptor(a.nulls), acu.host_descriptor(b.nulls)
        ab = (vp * max(len(a.buffers), 1))(*[x.ctypes.data for x in a.buffers])
        bb = (vp * max(len(b.buffers), 1))(*[x.ctypes.data for x in b.buffers])
        av, bv = np.ascontiguousarray(a.views), np.ascontiguousarray(b.views)
        self.check(self.lib.orc_cmp_byte_view(op, av.ctypes.data, ab, len(a.buffers), C.byref(ad), bv.ctypes.data, bb, len(b.buffers),
                                              C.byref(bd), C.byref(out)))
        return self._result(out, vals, valid, BOOL)

    # -- fused compare -> filter: by definition filter(values, cmp(a, b)) ------------------------
    def filter_cmp(self, values, op, a, b):
        pred = self.cmp(op, a, b)
        return self.filter(values, pred), self.filter_plan(pred)

    # -- nullif / zip --------------------------------------------------------------------------
    def nullif(self, left, right):
        n = max(left.length, 1)
        out, vals, valid = self._out(0, n)
        ld, rd = acu.host_descriptor(left), acu.host_descriptor(right)
        self.check(self.lib.orc_nullif(C.byref(ld), C.byref(rd), C.byref(out)))
        m = out.len
        if m == 0:
            return left
        validity = valid[: bitmap_bytes(m)].copy() if out.has_validity else None
        v = left.values if left.dtype == BOOL else left.values[:m]
        return HostArray(left.dtype, v, m, validity, 0, left.values_offset if left.dtype == BOOL else 0,
                         out.null_count if out.has_validity else 0)

    def zip(self, mask, truthy, falsy):
        n = mask.length
        out, vals, valid = self._out(n * truthy.width(), max(n, 1))
        md, td, fd = acu.host_descriptor(mask), acu.host_descriptor(truthy), acu.host_descriptor(falsy)
        self.check(self.lib.orc_zip(truthy.width(), C.byref(md), C.byref(td), C.byref(fd), C.byref(out)))
        return self._result(out, vals, valid, truthy.dtype)

    # -- boolean (arrow-arith/src/boolean.rs) -------------------------------------------------
    def boolean(self, op, a, b=None):
        n = max(a.length, 1)
        out, vals, valid = self._out(bitmap_bytes(n), n)
        ad = acu.host_descriptor(a)
        bd = acu.host_descriptor(b) if b is not None else None
        self.check(self.lib.orc_boolean(op, C.byref(ad), C.byref(bd) if bd is not None else None, C.byref(out)))
        return self._result(out, vals, valid, BOOL)

    def and_(self, a, b): return self.boolean(abi.BOOL_AND, a, b)
    def or_(self, a, b): return self.boolean(abi.BOOL_OR, a, b)
    def and_not(self, a, b): return self.boolean(abi.BOOL_AND_NOT, a, b)
    def and_kleene(self, a, b): return self.boolean(abi.BOOL_AND_KLEENE, a, b)
    def or_kleene(self, a, b): return self.boolean(abi.BOOL_OR_KLEENE, a, b)
    def not_(self, a): return self.boolean(abi.BOOL_NOT, a)
    def is_null(self, a): return self.boolean(abi.BOOL_IS_NULL, a)
    def is_not_null(self, a): return self.boolean(abi.BOOL_IS_NOT_NULL, a)

    def eq(self, a, b): return self.cmp(abi.EQ, a, b)
    def neq(self, a, b): return self.cmp(abi.NEQ, a, b)
    def lt(self, a, b): return self.cmp(abi.LT, a, b)
    def lt_eq(self, a, b): return self.cmp(abi.LT_EQ, a, b)
    def gt(self, a, b): return self.cmp(abi.GT, a, b)
    def gt_eq(self, a, b): return self.cmp(abi.GT_EQ, a, b)
    def distinct(self, a, b): return self.cmp(abi.DISTINCT, a, b)
    def not_distinct(self, a, b): return self.cmp(abi.NOT_DISTINCT, a, b)

    # -- cast / aggregate ----------------------------------------------------------------------
    def cast(self, a, to_dtype, safe=True):
        out, vals, valid = self._out(a.length * abi.DTYPE_SIZE[to_dtype], a.length)
        ad = acu.host_descriptor(a)
        self.check(self.lib.orc_cast_numeric(a.dtype, to_dtype, int(safe), C.byref(ad), C.byref(out)))
        return self._result(out, vals, valid, to_dtype)

    def aggregate(self, op, a, vector_bytes=16):
        bits, cnt = u64(0), i64(0)
        ad = acu.host_descriptor(a)
        self.check(self.lib.orc_aggregate(a.dtype, op, C.byref(ad), vector_bytes, C.byref(bits), C.byref(cnt)))
        if cnt.value == 0:
            return None
        raw = np.array([bits.value], dtype=np.uint64).view(np.uint8)[: abi.DTYPE_SIZE[a.dtype]]
        return raw.view(acu.NP_DTYPES[a.dtype])[0].item()

    def sum_checked(self, a):
        bits, cnt = u64(0), i64(0)
        ad = acu.host_descriptor(a)
        self.check(self.lib.orc_sum_checked(a.dtype, C.byref(ad), C.byref(bits), C.byref(cnt)))
        if cnt.value == 0:
            return None
        raw = np.array([bits.value], dtype=np.uint64).view(np.uint8)[: abi.DTYPE_SIZE[a.dtype]]
        return raw.view(acu.NP_DTYPES[a.dtype])[0].item()

    def sum(self, a, vector_bytes=16): return self.aggregate(abi.SUM, a, vector_bytes)
    def min(self, a): return self.aggregate(abi.MIN, a)
    def max(self, a): return self.aggregate(abi.MAX, a)

    # -- generators ------------------------------------------------------------------------------
    def generate_values(self, kind, seed, first_row, param, n, np_dtype):
        out = np.zeros(n, dtype=np_dtype)
        self.check(self.lib.orc_generate_values(kind, seed, first_row, param, out.ctypes.data, n))
        return out

    def generate_bits(self, seed, first_row, p, n):
        out = np.zeros(bitmap_bytes(n) + 8, dtype=np.uint8)
        self.check(self.lib.orc_generate_bits(seed, first_row, p, out.ctypes.data, n))
        return out


class RefBench:
    """The native multi-threaded harness of oracle/refbench.cpp: the hot-path step (filter -> take -> add -> sum) of
    the CPU restatement over a row-partitioned synthetic table, persistent pinned threads, timer inside C.
    seeds = (values i64, a, b, i64 validity, a validity, b validity, predicate)."""

    CHECK_KEYS = ("filter_rows", "filter_nulls", "filter_values_wsum", "take_nulls", "take_values_wsum", "add_nulls",
                  "add_bits_wsum", "sum_valid_rows")

    def __init__(self, rows, seeds, selectivity, null_density, threads=0, first_row=0, pin=True, oracle=None):
        self.orc = oracle or Oracle()
        self.h = vp()
        gen = C.c_double(0)
        arr = (u64 * 7)(*seeds)
        self.orc.check(self.orc.lib.orc_bench_create(rows, first_row, threads, arr, selectivity, null_density, int(pin),
                                                     C.byref(self.h), C.byref(gen)))
        self.rows, self.generate_seconds = rows, gen.value
        self.threads = self.orc.lib.orc_bench_threads(self.h)

    def step(self):
        """-> (seconds, sum_bits, valid_rows)"""
        s, bits, valid = C.c_double(0), u64(0), i64(0)
        self.orc.check(self.orc.lib.orc_bench_step(self.h, C.byref(s), C.byref(bits), C.byref(valid)))
        return s.value, bits.value, valid.value

    def check(self):
        out = (u64 * 8)()
        self.orc.check(self.orc.lib.orc_bench_check(self.h, out))
        return dict(zip(self.CHECK_KEYS, [int(x) for x in out]))

    def close(self):
        if self.h:
            self.orc.lib.orc_bench_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class OracleViewBackend:
    """The per-view work of acu.coalesce_views.InProgressByteViewArray done by the CPU oracle over numpy arrays (the same policy
    class then yields the reference's layout: views, buffer lengths / capacities, contents)."""

    def __init__(self, oracle):
        self.lib = oracle.lib
        self._oracle = oracle
        self.lib.orc_view_bytes_used.restype = i64
        self.lib.orc_view_bytes_used.argtypes = [vp, i64]
        self.lib.orc_view_copy_strings.restype = i64
        self.lib.orc_view_copy_strings.argtypes = [vp, i64, P(vp), C.c_uint32, vp, i64, vp]

    def upload(self, col):
        views = np.ascontiguousarray(col.views).reshape(-1).copy()
        caps = getattr(col, "buffer_capacities", None)
        bufs = [(np.concatenate([b, np.zeros(16, np.uint8)]), int(b.nbytes), int(caps[i]) if caps else int(b.nbytes)) for i, b in enumerate(col.buffers)]
        return {"views": views, "n": col.length, "buffers": bufs, "refs": 0}

    def release_source(self, src):
        pass

    def bytes_used(self, src):
        return self.lib.orc_view_bytes_used(src["views"].ctypes.data, src["n"])

    def fit(self, src, offset, n, remaining):
        nv, nb = i64(0), i64(0)
        self.lib.orc_view_fit(src["views"].ctypes.data + 16 * offset, n, remaining, C.byref(nv), C.byref(nb))
        return nv.value, nb.value

    def copy_strings(self, src, offset, n, new_index, dst, dst_len, dst_cap, out_views, out_at):
        table = (vp * max(len(src["buffers"]), 1))(*[b.ctypes.data for b, _, _ in src["buffers"]])
        nb = self.lib.orc_view_copy_strings(src["views"].ctypes.data + 16 * offset, n, table, new_index, dst.ctypes.data, dst_len,
                                            out_views.ctypes.data + 16 * out_at)
        assert dst_len + nb <= dst_cap
        return nb

    def rebase(self, src, offset, n, delta, out_views, out_at):
        self.lib.orc_view_rebase(src["views"].ctypes.data + 16 * offset, n, delta, out_views.ctypes.data + 16 * out_at)

    def filter_views(self, col, predicate):
        """filter_byte_view = filter_native(views) + filter_nulls, buffers shared (filter.rs:931-944), through the oracle's
        16-byte-wide filter."""
        o = self._oracle
        n = col.length
        views = np.ascontiguousarray(col.views).reshape(-1).copy()
        vd = abi.Array()
        vd.values, vd.values_offset = views.ctypes.data, 0
        vd.validity = col.nulls.validity.ctypes.data if col.nulls.validity is not None else None
        vd.validity_offset, vd.len, vd.is_scalar = col.nulls.validity_offset, n, 0
        vd.null_count = col.nulls.null_count if col.nulls.validity is not None else 0
        out, vals, valid = o._out(max(n, predicate.length) * 16, max(n, predicate.length))
        pd = acu.host_descriptor(predicate)
        cnt, strat = i64(0), i32(0)
        o.check(o.lib.orc_filter_primitive(C.byref(pd), 16, C.byref(vd), C.byref(out), C.byref(cnt), C.byref(strat)))
        m = out.len
        validity = valid[: bitmap_bytes(m)].copy() if out.has_validity else None
        res = acu.ViewColumn(vals[: m * 16].copy().reshape(m, 16), col.buffers,
                             HostArray(acu.U8, np.zeros(0, np.uint8), m, validity, 0, 0, out.null_count if out.has_validity else 0))
        res.buffer_capacities = getattr(col, "buffer_capacities", None)
        return res

    def alloc(self, nbytes):
        return np.zeros(max(nbytes, 16) + 16, dtype=np.uint8)

    def free(self, p):
        pass

    def download(self, p, nbytes):
        return p[:nbytes].copy()
