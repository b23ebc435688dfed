"""Utf8View coalescing policy (acu/coalesce_views.py = InProgressByteViewArray, arrow-select/src/coalesce/byte_view.rs) driven by
the CPU oracle's per-view functions, pinned on the reference's own tests: BufferSource sizes (byte_view.rs:568-606) and the
expected data-buffer layouts of test_string_view_* (arrow-select/src/coalesce.rs:1046-1420)."""
import numpy as np
import pytest

from acu.coalesce_views import BufferSource, ViewBatchCoalescer
from oracle import OracleViewBackend
from view_util import as_bytes, view_batch, view_slice, view_values

LONG = "This string is longer than 12 bytes"


def run(oracle, batches, batch_size):
    co = ViewBatchCoalescer(OracleViewBackend(oracle), batch_size)
    expect = []
    for b in batches:
        expect += view_values(b)
        co.push_batch(b)
    co.finish_buffered_batch()
    got = []
    for col, _ in co.completed:
        got += view_values(col)
    assert got == expect  # the output is the concatenation of the inputs (coalesce.rs:84-146)
    return co.completed


def test_buffer_source():
    s = BufferSource()  # byte_view.rs:568-583
    assert [s.next_size(1000) for _ in range(9)] == [8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576, 1048576]
    assert s.next_size(10_000_000) == 10_000_000
    s = BufferSource()  # :585-594
    assert [s.next_size(5_600) for _ in range(3)] == [8192, 16384, 32768]
    s = BufferSource()  # :596-606
    assert [s.next_size(500_000) for _ in range(3)] == [512 * 1024, 1024 * 1024, 1024 * 1024]
    assert s.next_size(2_000_000) == 2_000_000


def test_string_view_no_views(oracle):  # coalesce.rs:1046-1058
    out = run(oracle, [view_batch(2, ["foo", "bar"]), view_batch(2, ["baz", "qux"])], 1024)
    assert [c.length for c, _ in out] == [4] and out[0][1] == []


def test_string_view_batch_small_no_compact(oracle):  # :1061-1076
    out = run(oracle, [view_batch(1000, ["a", "b", "c"])], 1024)
    assert [c.length for c, _ in out] == [1000] and out[0][1] == []


def test_string_view_batch_large_no_compact(oracle):  # :1079-1118
    b = view_batch(1000, [LONG])
    assert len(b.buffers) == 5
    out = run(oracle, [b], 1000)
    assert out[0][1] == [(8190, 8192)] * 4 + [(2240, 8192)]


def test_string_view_batch_small_with_buffers_no_compact(oracle):  # :1121-1140
    b = view_slice(view_batch(1000, ["SmallString"] * 20 + [LONG]), 5, 10)
    out = run(oracle, [b], 1000)
    assert [c.length for c, _ in out] == [10] and out[0][1] == []


def test_string_view_batch_large_slice_compact(oracle):  # :1143-1167
    out = run(oracle, [view_slice(view_batch(1000, [LONG]), 11, 22)], 1000)
    assert [c.length for c, _ in out] == [22] and out[0][1] == [(770, 8192)]


def test_string_view_mixed(oracle):  # :1170-1227
    large, small = view_batch(1000, [LONG]), view_batch(1000, ["SmallString"])
    mixed, mixed_nulls = view_batch(1000, [LONG, "Small"]), view_batch(1000, [LONG, "Small", None])
    out = run(oracle, [large, small, view_slice(large, 10, 20), mixed_nulls, view_slice(large, 10, 20), mixed], 1024)
    assert [c.length for c, _ in out] == [1024, 1024, 1024, 968]
    assert out[0][1] == [(8190, 8192)] * 4 + [(2240, 8192)]


def test_string_view_many_small_compact(oracle):  # :1230-1273
    b = view_batch(200, ["This string is 28 bytes long", "small string"])
    out = run(oracle, [b] * 10, 8000)
    assert [c.length for c, _ in out] == [2000]
    assert out[0][1] == [(8176, 8192), (16380, 16384), (3444, 32768)]


def test_string_view_many_small_boundary(oracle):  # :1276-1304
    b = view_batch(100, ["This string is a power of two=32"])
    out = run(oracle, [b] * 20, 900)
    assert [c.length for c, _ in out] == [900, 900, 200]
    assert out[0][1] == [(8192, 8192), (16384, 16384), (4224, 32768)]


def test_fuzz_content_and_nulls(oracle):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
    batches = []
    for _ in range(12):
        n = int(rng.integers(1, 700))
        items = []
        for _ in range(n):
            r = rng.random()
            if r < 0.1:
                items.append(None)
            else:
                ln = int(rng.integers(0, 13)) if r < 0.5 else int(rng.integers(13, 300))
                items.append(bytes(rng.choice(alphabet, ln)))
        b = view_batch(n, items, int(rng.choice([64, 512, 8192])))
        if rng.random() < 0.5 and n > 4:
            off = int(rng.integers(0, n // 2))
            b = view_slice(b, off, int(rng.integers(1, n - off)))
        batches.append(b)
    out = run(oracle, batches, 257)
    assert all(c.length == 257 for c, _ in out[:-1])


def bool_array(mask):
    from acu import HostArray
    return HostArray.bool_from_numpy(np.asarray(mask, dtype=bool))


def run_filtered(oracle, steps, batch_size):
    """steps: [(ViewColumn, predicate mask or None)] -> completed [(ViewColumn, layout)], checking the logical content."""
    co = ViewBatchCoalescer(OracleViewBackend(oracle), batch_size)
    expect = []
    for b, mask in steps:
        vals = view_values(b)
        if mask is None:
            expect += vals
            co.push_batch(b)
        else:
            expect += [v for v, m in zip(vals, mask) if m]
            co.push_batch_with_filter(b, bool_array(mask))
    co.finish_buffered_batch()
    got = []
    for col, _ in co.completed:
        got += view_values(col)
    assert got == expect
    return co.completed


def test_string_view_filtered_inline(oracle):  # coalesce.rs:1424-1441 (+ :1404-1421 BinaryView): idx % 8 == 0 of 1000 rows, twice
    b = view_batch(1000, ["foo", None, "barbaz"])
    mask = [i % 8 == 0 for i in range(1000)]
    out = run_filtered(oracle, [(b, mask), (b, mask)], 300)
    assert [c.length for c, _ in out] == [250] and out[0][1] == []


def test_inline_view_very_sparse(oracle):  # coalesce.rs:1503-1520: idx % 20 == 0 (the sparse per-column copy path)
    b = view_batch(1000, ["foo", None, "barbaz"])
    mask = [i % 20 == 0 for i in range(1000)]
    out = run_filtered(oracle, [(b, mask), (b, mask)], 1024)
    assert [c.length for c, _ in out] == [100] and out[0][1] == []


def test_copy_rows_by_filter_from_reuses_non_inline_buffers(oracle):  # byte_view.rs:619-650
    vals = [f"This value is longer than 12 bytes: {i}" for i in range(32)]
    b = view_batch(32, vals)
    assert len(b.buffers) == 1
    mask = [i == 3 or i == 29 for i in range(32)]  # 2 of 32 rows: sparse path (2 <= 32 / 16)
    out = run_filtered(oracle, [(b, mask)], 32)
    col, layout = out[0]
    assert view_values(col) == as_bytes([vals[3], vals[29]])
    assert layout == [(int(b.buffers[0].nbytes), 8192)]  # the source's buffer, adopted as it is (no copy, no compaction)
    assert np.array_equal(col.buffers[0], b.buffers[0])


def test_filtered_dense_goes_through_push_batch(oracle):
    """Not sparse enough (selected > len / 16): filter_record_batch, then push_batch — the gc decision is taken on the filtered
    array (coalesce.rs:652-666): 125 of 1000 long strings use 4375 of 40960 buffer bytes => compacted into one 8 KiB buffer."""
    b = view_batch(1000, [LONG])
    mask = [i % 8 == 0 for i in range(1000)]
    out = run_filtered(oracle, [(b, mask)], 1000)
    assert [c.length for c, _ in out] == [125] and out[0][1] == [(125 * 35, 8192)]
