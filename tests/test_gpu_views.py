"""Utf8View / BinaryView coalescing on the device (csrc/views.cu behind acu/coalesce_views.py) against the CPU oracle: the SAME
policy class (InProgressByteViewArray, arrow-select/src/coalesce/byte_view.rs) runs once over the device entry points and once
over the oracle's per-view functions; output views, data buffers (bytes, lengths, capacities) and nulls must be identical,
and equal to the reference's expected layouts (coalesce.rs:1046-1420, pinned on the oracle in tests/test_oracle_views.py)."""
import ctypes as C

import numpy as np
import pytest

from acu import _abi as abi
from acu.coalesce_views import DeviceViewBackend, ViewBatchCoalescer
from oracle import OracleViewBackend
from view_util import view_batch, view_slice, view_values

pytestmark = pytest.mark.gpu

LONG = "This string is longer than 12 bytes"


def both(gpu, oracle, batches, batch_size):
    outs = []
    for be in (DeviceViewBackend(gpu), OracleViewBackend(oracle)):
        co = ViewBatchCoalescer(be, batch_size)
        for b in batches:
            co.push_batch(b)
        co.finish_buffered_batch()
        outs.append(co.completed)
        co.close()
    g, e = outs
    assert len(g) == len(e)
    for (gc, gl), (ec, el) in zip(g, e):
        assert gl == el, "buffer layout (len, capacity)"
        assert np.array_equal(gc.views, ec.views), "views"
        assert len(gc.buffers) == len(ec.buffers) and all(np.array_equal(x, y) for x, y in zip(gc.buffers, ec.buffers)), "data buffers"
        assert (gc.nulls.validity is None) == (ec.nulls.validity is None) and gc.nulls.null_count == ec.nulls.null_count
        assert view_values(gc) == view_values(ec)
    return g


def test_reference_layouts(gpu, oracle):
    large = view_batch(1000, [LONG])
    assert both(gpu, oracle, [large], 1000)[0][1] == [(8190, 8192)] * 4 + [(2240, 8192)]            # coalesce.rs:1079-1118
    assert both(gpu, oracle, [view_slice(large, 11, 22)], 1000)[0][1] == [(770, 8192)]              # :1143-1167
    b = view_batch(200, ["This string is 28 bytes long", "small string"])
    assert both(gpu, oracle, [b] * 10, 8000)[0][1] == [(8176, 8192), (16380, 16384), (3444, 32768)]  # :1230-1273
    b = view_batch(100, ["This string is a power of two=32"])
    out = both(gpu, oracle, [b] * 20, 900)                                                           # :1276-1304
    assert [c.length for c, _ in out] == [900, 900, 200] and out[0][1] == [(8192, 8192), (16384, 16384), (4224, 32768)]
    small = view_batch(1000, ["SmallString"])
    mixed, mixed_nulls = view_batch(1000, [LONG, "Small"]), view_batch(1000, [LONG, "Small", None])
    out = both(gpu, oracle, [large, small, view_slice(large, 10, 20), mixed_nulls, view_slice(large, 10, 20), mixed], 1024)  # :1170-1227
    assert [c.length for c, _ in out] == [1024, 1024, 1024, 968]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz(gpu, oracle, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
    batches = []
    for _ in range(10):
        n = int(rng.integers(1, 3000))
        items = []
        for _ in range(n):
            r = rng.random()
            if r < 0.1:
                items.append(None)
            else:
                ln = int(rng.integers(0, 13)) if r < 0.5 else int(rng.integers(13, 100)) if r < 0.97 else int(rng.integers(200, 5000))
                items.append(bytes(rng.choice(alphabet, ln)))
        b = view_batch(n, items, int(rng.choice([64, 4096, 8192])))
        if rng.random() < 0.6 and n > 4:
            off = int(rng.integers(0, n // 2))
            b = view_slice(b, off, int(rng.integers(1, n - off)))
        batches.append(b)
    both(gpu, oracle, batches, int(rng.choice([100, 1024, 4000])))


def test_primitives_edges(gpu, oracle):
    lib, h = gpu.lib, gpu.h
    col = view_batch(300, [LONG, "tiny", "exactly12byt", "thirteen byte"])
    be_g, be_o = DeviceViewBackend(gpu), OracleViewBackend(oracle)
    dg, do = be_g.upload(col), be_o.upload(col)
    assert be_g.bytes_used(dg) == be_o.bytes_used(do) == 75 * (35 + 13)
    for remaining in (0, 3, 4, 12, 13, 34, 35, 47, 48, 49, 1000, 10**9):
        for off, n in ((0, 300), (1, 299), (2, 17), (3, 1)):
            assert be_g.fit(dg, off, n, remaining) == be_o.fit(do, off, n, remaining), (remaining, off, n)
    # a destination that is too small, and a view pointing outside its array's buffers, are reported, not executed
    out_views, dst = gpu.malloc(300 * 16 + 16), gpu.malloc(64)
    table = (C.c_void_p * 1)(dg["buffers"][0][0])
    nb = C.c_int64(0)
    st = lib.acu_view_copy_strings(h, dg["views"], 300, table, 1, 0, dst, 0, 32, out_views, C.byref(nb))
    assert st == abi.ERR_INVALID_ARGUMENT and b"do not fit" in lib.acu_last_error(h).contents.message
    big = gpu.malloc(75 * 48 + 64)
    st = lib.acu_view_copy_strings(h, dg["views"], 300, table, 0, 0, big, 0, 75 * 48, out_views, C.byref(nb))
    assert st == abi.ERR_INVALID_ARGUMENT and b"refers to a data buffer" in lib.acu_last_error(h).contents.message
    assert lib.acu_view_copy_strings(h, dg["views"], 0, table, 1, 0, big, 0, 0, out_views, C.byref(nb)) == abi.OK and nb.value == 0
    for p in (out_views, dst, big):
        gpu.free(p)
    be_g.release_source(dg)


def both_filtered(gpu, oracle, steps, batch_size):
    from acu import HostArray
    outs = []
    for be in (DeviceViewBackend(gpu), OracleViewBackend(oracle)):
        co = ViewBatchCoalescer(be, batch_size)
        for b, mask in steps:
            if mask is None:
                co.push_batch(b)
            else:
                co.push_batch_with_filter(b, HostArray.bool_from_numpy(np.asarray(mask, dtype=bool)))
        co.finish_buffered_batch()
        outs.append(co.completed)
        co.close()
    g, e = outs
    assert len(g) == len(e)
    for (gc, gl), (ec, el) in zip(g, e):
        assert gl == el and np.array_equal(gc.views, ec.views)
        assert len(gc.buffers) == len(ec.buffers) and all(np.array_equal(x, y) for x, y in zip(gc.buffers, ec.buffers))
        assert (gc.nulls.validity is None) == (ec.nulls.validity is None) and gc.nulls.null_count == ec.nulls.null_count
        assert view_values(gc) == view_values(ec)
    return g


def test_push_batch_with_filter(gpu, oracle):
    """push_batch_with_filtered_columns (coalesce.rs:620-681) for view columns: the sparse per-column copy (views filtered on the
    device by the 16-byte filter kernel, source buffers adopted) and the materialised path (filter, then push_batch with its gc
    decision) — identical layouts on the device and on the oracle; reference cases coalesce.rs:1424-1441,1503-1520,
    byte_view.rs:619-650."""
    inline = view_batch(1000, ["foo", None, "barbaz"])
    out = both_filtered(gpu, oracle, [(inline, [i % 8 == 0 for i in range(1000)])] * 2, 300)
    assert [c.length for c, _ in out] == [250]
    out = both_filtered(gpu, oracle, [(inline, [i % 20 == 0 for i in range(1000)])] * 2, 1024)
    assert [c.length for c, _ in out] == [100]
    vals = [f"This value is longer than 12 bytes: {i}" for i in range(32)]
    b = view_batch(32, vals)
    out = both_filtered(gpu, oracle, [(b, [i == 3 or i == 29 for i in range(32)])], 32)
    assert out[0][1] == [(int(b.buffers[0].nbytes), 8192)]
    large = view_batch(1000, [LONG])
    out = both_filtered(gpu, oracle, [(large, [i % 8 == 0 for i in range(1000)])], 1000)
    assert out[0][1] == [(125 * 35, 8192)]
    # mixed sequence with nulls, slices, dense and sparse filters across batch boundaries
    rng = np.random.default_rng(5)
    mixed = view_batch(3000, [LONG, "Small", None, "another rather long string value", "x"])
    steps = [(mixed, None), (view_slice(mixed, 100, 2000), rng.random(2000) < 0.03), (mixed, rng.random(3000) < 0.5),
             (view_slice(mixed, 7, 900), rng.random(900) < 0.05), (large, rng.random(1000) < 0.04)]
    both_filtered(gpu, oracle, steps, 700)
