"""Utf8View / BinaryView columns in BatchCoalescer: the host mirror of InProgressByteViewArray
(arrow-select/src/coalesce/byte_view.rs:39-520) over the device entry points of csrc/views.cu.

The POLICY is the reference's, line by line: a source array whose data buffers hold more than twice the bytes its views use
is garbage-collected (its long strings are copied into the coalescer's own buffers, :366-381), otherwise its buffers are
adopted and the views' buffer indices rebased (:176-216); output buffers come from BufferSource (8 KiB doubling to 1 MiB, or
the size asked for if larger, :526-559); a buffer that cannot take all of a source's strings is filled with the views that
fit and a new one is started (:246-291). The per-view WORK (total_buffer_bytes_used, the fit loop, the copy + rewrite, the
rebase) is the backend's: `DeviceViewBackend` (acu_view_* on device memory) here, an oracle backend over numpy in the tests.
"""
import ctypes as C

import numpy as np

from . import U8, HostArray, ViewColumn, bitmap_bytes, pack_bits

STARTING_BLOCK_SIZE = 4 * 1024  # (the first size handed out is 8 KiB: byte_view.rs:521)
MAX_BLOCK_SIZE = 1024 * 1024


class BufferSource:
    """byte_view.rs:526-559."""

    def __init__(self):
        self.current_size = STARTING_BLOCK_SIZE

    def next_size(self, min_size):
        if self.current_size < MAX_BLOCK_SIZE:
            self.current_size *= 2
        if self.current_size >= min_size:
            return self.current_size
        while self.current_size <= min_size and self.current_size < MAX_BLOCK_SIZE:
            self.current_size *= 2
        return max(self.current_size, min_size)


class DeviceViewBackend:
    """Views and data buffers in HBM; every method is one acu_view_* call (+ allocation / plain copies)."""

    def __init__(self, ctx):
        self.ctx = ctx

    # -- sources ------------------------------------------------------------------------------------
    def upload(self, col):
        ctx = self.ctx
        n = col.length
        views = np.ascontiguousarray(col.views).reshape(-1)
        d_views = ctx.malloc(max(views.nbytes, 16) + 16)
        if views.nbytes:
            ctx.h2d(d_views, views)
        bufs = []
        caps = getattr(col, "buffer_capacities", None)  # Buffer::capacity() of each data buffer (default: its length)
        for i, b in enumerate(col.buffers):
            d = ctx.malloc(max(b.nbytes, 1) + 16)
            if b.nbytes:
                ctx.h2d(d, b)
            bufs.append((d, int(b.nbytes), int(caps[i]) if caps else int(b.nbytes)))
        return {"views": d_views, "n": n, "buffers": bufs, "refs": 0}

    def release_source(self, src):
        """Called when the last reference (current source / in-progress batches that adopted its buffers) is gone."""
        self.ctx.free(src["views"])
        for d, _, _ in src["buffers"]:
            self.ctx.free(d)

    # -- per-view work --------------------------------------------------------------------------------
    def bytes_used(self, src):
        out = C.c_int64(0)
        self.ctx.check(self.ctx.lib.acu_view_bytes_used(self.ctx.h, src["views"], src["n"], C.byref(out)))
        return out.value

    def fit(self, src, offset, n, remaining):
        nv, nb = C.c_int64(0), C.c_int64(0)
        self.ctx.check(self.ctx.lib.acu_view_fit(self.ctx.h, src["views"] + 16 * offset, n, remaining, C.byref(nv), C.byref(nb)))
        return nv.value, nb.value

    def copy_strings(self, src, offset, n, new_index, dst, dst_len, dst_cap, out_views, out_at):
        table = (C.c_void_p * max(len(src["buffers"]), 1))(*[d for d, _, _ in src["buffers"]])
        nb = C.c_int64(0)
        self.ctx.check(self.ctx.lib.acu_view_copy_strings(self.ctx.h, src["views"] + 16 * offset, n, table, len(src["buffers"]), new_index, dst,
                                                          dst_len, dst_cap, out_views + 16 * out_at, C.byref(nb)))
        return nb.value

    def rebase(self, src, offset, n, delta, out_views, out_at):
        self.ctx.check(self.ctx.lib.acu_view_rebase(self.ctx.h, src["views"] + 16 * offset, n, delta, out_views + 16 * out_at))

    def filter_views(self, col, predicate):
        """filter_byte_view (arrow-select/src/filter.rs:931-944): filter_native over the 16-byte views + filter_nulls; the data
        buffers are shared. col: ViewColumn, predicate: HostArray(BOOL) -> ViewColumn (host)."""
        from . import _abi as abi
        ctx = self.ctx
        n = col.length
        views = np.ascontiguousarray(col.views).reshape(-1)
        d_views = ctx.malloc(max(views.nbytes, 16) + 16)
        if views.nbytes:
            ctx.h2d(d_views, views)
        d_valid = None
        if col.nulls.validity is not None:
            d_valid = ctx.malloc(col.nulls.validity.nbytes + 8)
            ctx.h2d(d_valid, col.nulls.validity)
        vd = abi.Array()
        vd.values, vd.values_offset, vd.validity, vd.validity_offset = d_views, 0, d_valid, col.nulls.validity_offset
        vd.len, vd.null_count, vd.is_scalar = n, (col.nulls.null_count if d_valid else 0), 0
        dp = ctx.upload(predicate)
        plan = C.c_void_p()
        out = None
        try:
            pd = dp.descriptor()
            ctx.check(ctx.lib.acu_filter_plan_create(ctx.h, C.byref(pd), C.byref(plan)))
            count = ctx.lib.acu_filter_plan_count(plan)
            out = ctx.alloc_out(count * 16, count)
            ctx.check(ctx.lib.acu_filter_primitive(ctx.h, plan, 16, C.byref(vd), C.byref(out)))
            fv = ctx.d2h(out.values, count * 16).reshape(count, 16) if count else np.zeros((0, 16), np.uint8)
            validity = ctx.d2h(out.validity, bitmap_bytes(count)) if out.has_validity else None
            nulls = HostArray(U8, np.zeros(0, np.uint8), count, validity, 0, 0, out.null_count if out.has_validity else 0)
        finally:
            if out is not None:
                ctx._free_out(out)
            if plan:
                ctx.lib.acu_filter_plan_destroy(ctx.h, plan)
            dp.free()
            ctx.free(d_views)
            if d_valid:
                ctx.free(d_valid)
        res = ViewColumn(fv, col.buffers, nulls)
        res.buffer_capacities = getattr(col, "buffer_capacities", None)
        return res

    # -- storage ----------------------------------------------------------------------------------------
    def alloc(self, nbytes):
        return self.ctx.malloc(max(nbytes, 16) + 16)

    def free(self, p):
        self.ctx.free(p)

    def download(self, p, nbytes):
        return self.ctx.d2h(p, nbytes) if nbytes else np.zeros(0, np.uint8)


class InProgressByteViewArray:
    """byte_view.rs:39-520 (push path: set_source / copy_rows / finish). `backend` does the per-view work."""

    def __init__(self, backend, batch_size):
        self.be, self.batch_size = backend, batch_size
        self.buffer_source = BufferSource()
        self.source = None
        self._reset()

    def _reset(self):
        self.views, self.n_views = None, 0        # output views (allocated on first write: ensure_capacity)
        self.valid = []                           # per-piece validity (host bools): NullBufferBuilder
        self.current = None                       # [buffer, len, capacity]
        self.completed = []                       # [(buffer, len, capacity, owned)]
        self.adopted_from = []                    # sources whose buffers this in-progress batch shares (the reference: Arc)

    # -- set_source (:357-391) --------------------------------------------------------------------------
    def set_source(self, col):
        """col: ViewColumn (host). Uploaded through the backend; gc decision as the reference's."""
        if self.source is not None:
            self._unref(self.source["dev"])
        if col is None:
            self.source = None
            return
        dev = self.be.upload(col)
        dev["refs"] = 1
        if not dev["buffers"]:
            need_gc, ideal = False, 0
        else:
            ideal = self.be.bytes_used(dev)
            actual = sum(cap for _, _, cap in dev["buffers"])
            need_gc = ideal != 0 and actual > ideal * 2
        self.source = {"col": col, "dev": dev, "need_gc": need_gc, "ideal": ideal}

    def _unref(self, dev):
        dev["refs"] -= 1
        if dev["refs"] == 0:
            self.be.release_source(dev)

    # -- copy_rows (:393-436) ---------------------------------------------------------------------------
    def copy_rows(self, offset, n):
        if self.views is None:
            self.views = self.be.alloc(self.batch_size * 16)
        s = self.source
        col = s["col"]
        if col.nulls.validity is not None:
            self.valid.append(col.nulls.valid_mask()[offset:offset + n].copy())
        else:
            self.valid.append(np.ones(n, dtype=bool))
        dev = s["dev"]
        if s["ideal"] == 0:  # all views inline (or no buffers): views are appended as they are
            self.be.rebase(dev, offset, n, 0, self.views, self.n_views)
        elif s["need_gc"]:
            self._append_views_and_copy_strings(dev, offset, n, s["ideal"])
        else:
            self._append_views_and_update_buffer_index(dev, offset, n, s)
        self.n_views += n

    def copy_rows_by_filter_from(self, source_col, filtered_col):
        """byte_view.rs:464-488: the sparse-filter path. All-inline sources: the filtered views / nulls are appended as they
        are; sources with data buffers: the filtered views keep pointing into the SOURCE's buffers, which are adopted
        (append_views_and_update_buffer_index(.., is_reused = false)) — no string is copied, no gc decision is taken."""
        if self.views is None:
            self.views = self.be.alloc(self.batch_size * 16)
        n = filtered_col.length
        if filtered_col.nulls.validity is not None:
            self.valid.append(filtered_col.nulls.valid_mask()[:n].copy())
        else:
            self.valid.append(np.ones(n, dtype=bool))
        dev = self.be.upload(filtered_col)
        dev["refs"] = 1
        if not source_col.buffers:
            self.be.rebase(dev, 0, n, 0, self.views, self.n_views)
        else:
            self._append_views_and_update_buffer_index(dev, 0, n, None)
        self._unref(dev)  # (an adopting batch holds its own reference)
        self.n_views += n

    def _finish_current(self):
        if self.current is not None:
            self.completed.append((self.current[0], self.current[1], self.current[2], True))
            self.current = None

    def _append_views_and_update_buffer_index(self, dev, offset, n, s):  # :176-216
        self._finish_current()
        starting = len(self.completed)
        for d, ln, cap in dev["buffers"]:
            self.completed.append((d, ln, cap, False))  # adopted: the source's buffers, shared
        if not any(x is dev for x in self.adopted_from):
            dev["refs"] += 1
            self.adopted_from.append(dev)
        self.be.rebase(dev, offset, n, starting, self.views, self.n_views)

    def _append_views_and_copy_strings(self, dev, offset, n, view_buffer_size):  # :228-291
        if self.current is None:
            cap = self.buffer_source.next_size(view_buffer_size)
            self._copy_inner(dev, offset, n, [self.be.alloc(cap), 0, cap], 0)
            return
        remaining = self.current[2] - self.current[1]
        if view_buffer_size <= remaining:
            cur, self.current = self.current, None
            self._copy_inner(dev, offset, n, cur, 0)
            return
        num_to_current, bytes_to_current = self.be.fit(dev, offset, n, remaining)
        remaining_view_buffer_size = view_buffer_size - bytes_to_current
        cur, self.current = self.current, None
        self._copy_inner(dev, offset, num_to_current, cur, 0)
        self._finish_current()
        cap = self.buffer_source.next_size(remaining_view_buffer_size)
        self._copy_inner(dev, offset + num_to_current, n - num_to_current, [self.be.alloc(cap), 0, cap], num_to_current)

    def _copy_inner(self, dev, offset, n, dst, out_skip):  # :298-354
        assert self.current is None
        if n == 0:
            self.current = dst
            return
        new_index = len(self.completed)
        nb = self.be.copy_strings(dev, offset, n, new_index, dst[0], dst[1], dst[2], self.views, self.n_views + out_skip)
        dst[1] += nb
        self.current = dst

    # -- finish (:490-520) ----------------------------------------------------------------------------------
    def finish(self):
        """-> (ViewColumn on the host, [(len, capacity)] of its data buffers); resets the in-progress state."""
        self._finish_current()
        n = self.n_views
        views = self.be.download(self.views, n * 16).reshape(n, 16) if n else np.zeros((0, 16), np.uint8)
        buffers, layout = [], []
        for buf, ln, cap, owned in self.completed:
            buffers.append(self.be.download(buf, ln))
            layout.append((ln, cap))
        valid = np.concatenate(self.valid) if self.valid else np.zeros(0, dtype=bool)
        if valid.all():
            nulls = HostArray(U8, np.zeros(0, np.uint8), n, None, 0, 0, 0)
        else:
            nulls = HostArray(U8, np.zeros(0, np.uint8), n, pack_bits(valid), 0, 0, int(n - valid.sum()))
        out = ViewColumn(views, buffers, nulls)
        for buf, ln, cap, owned in self.completed:
            if owned:
                self.be.free(buf)
        for dev in self.adopted_from:  # shared source buffers: released with their last user
            self._unref(dev)
        if self.views is not None:
            self.be.free(self.views)
        self._reset()
        return out, layout

    def close(self):
        self.set_source(None)
        if self.views is not None:
            self.finish()


class ViewBatchCoalescer:
    """BatchCoalescer::push_batch (coalesce.rs:488-529) for ONE Utf8View / BinaryView column."""

    def __init__(self, backend, target_batch_size):
        self.target = target_batch_size
        self.col = InProgressByteViewArray(backend, target_batch_size)
        self.buffered = 0
        self.completed = []

    def push_batch(self, view_column):
        n, offset = view_column.length, 0
        self.col.set_source(view_column)
        while n > self.target - self.buffered:
            remaining = self.target - self.buffered
            self.col.copy_rows(offset, remaining)
            self.buffered += remaining
            offset += remaining
            n -= remaining
            self.finish_buffered_batch()
        if n > 0:
            self.col.copy_rows(offset, n)
        self.buffered += n
        if self.buffered >= self.target:
            self.finish_buffered_batch()
        # the reference drops the source here (set_source(None), coalesce.rs:524-527); adopted buffers live on by refcount,
        # here they stay with the in-progress array until the next batch arrives

    def push_batch_with_filter(self, view_column, predicate):
        """push_batch_with_filtered_columns (coalesce.rs:620-681). predicate: HostArray(BOOL), no longer than the column."""
        from . import ArrowError
        from . import _abi as abi
        n = view_column.length
        if predicate.length > n:
            raise ArrowError(abi.ERR_INVALID_ARGUMENT,
                             f"Invalid argument error: Filter predicate of length {predicate.length} is larger than target array of length {n}")
        sel = predicate.to_numpy_bool() if hasattr(predicate, "to_numpy_bool") else None
        if sel is None:
            bits = np.unpackbits(np.asarray(predicate.values, dtype=np.uint8), bitorder="little")[predicate.values_offset:predicate.values_offset + predicate.length].astype(bool)
            if predicate.validity is not None:
                bits &= np.unpackbits(predicate.validity, bitorder="little")[predicate.validity_offset:predicate.validity_offset + predicate.length].astype(bool)
            sel = bits
        selected = int(sel.sum())
        if selected == 0:
            return
        if selected == n and predicate.length == n:
            return self.push_batch(view_column)
        does_not_fit = selected > self.target - self.buffered
        sparse_ok = selected <= predicate.length // 16  # should_use_sparse_filter_copy (coalesce.rs:49-54)
        filtered = self.col.be.filter_views(view_column, predicate)
        if does_not_fit or not sparse_ok:
            return self.push_batch(filtered)  # materialised filter, then the normal path (gc decision on the filtered array)
        self.col.copy_rows_by_filter_from(view_column, filtered)
        self.buffered += selected
        if self.buffered >= self.target:
            self.finish_buffered_batch()

    def finish_buffered_batch(self):
        if self.buffered == 0:
            return
        self.completed.append(self.col.finish())
        self.buffered = 0

    def close(self):
        self.col.close()
