"""BatchCoalescer on device (acu/coalesce.py over acu_bitmap_copy / acu_bitmap_fill /
acu_offsets_append / acu_memcpy_d2d and the record-batch filter / take) against the reference's
definition: the output is the concatenation of the pushed (filtered / taken) rows, cut into
batches of exactly target_batch_size rows in input order, the tail produced by
finish_buffered_batch (arrow-select/src/coalesce.rs:84-146). Literal cases transcribe the
reference's doc tests (coalesce.rs:42-110, :239-257, :271-288, :312-324) and
test_coalesce* / test_coalesce_filtered_* (:794-1044); the fuzz builds the expectation with the
oracle's filter / take column by column."""
import ctypes as C

import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray, Utf8Column
from acu.coalesce import BatchCoalescer
from golden_util import strings_of
from test_gpu_parity import rand_array, rand_bool, rand_strings

pytestmark = pytest.mark.gpu


def logical(col):
    if isinstance(col, Utf8Column):
        return strings_of(col.offsets, col.data, col.nulls)
    return col.to_list()


def has_nulls_buffer(col):
    return (col.nulls if isinstance(col, Utf8Column) else col).validity is not None


def drain(co):
    out = []
    while co.has_completed_batch():
        out.append(co.next_completed_batch())
    return out


def i32(items):
    return HostArray.from_list(abi.I32, items)


def test_doc_examples(gpu):
    # coalesce.rs:42-110: two 3-row batches into 4-row outputs
    co = BatchCoalescer(gpu, [abi.I32], 4)
    co.push_batch([i32([1, 2, 3])])
    assert co.next_completed_batch() is None and co.get_buffered_rows() == 3
    co.push_batch([i32([4, 5, 6])])
    b = co.next_completed_batch()
    assert logical(b[0]) == [1, 2, 3, 4] and co.next_completed_batch() is None
    co.finish_buffered_batch()
    assert logical(co.next_completed_batch()[0]) == [5, 6] and co.is_empty()
    co.close()
    # coalesce.rs:239-257 push_batch_with_filter
    co = BatchCoalescer(gpu, [abi.I32], 1000)
    f = HostArray.bool_from_numpy(np.array([True, False, True]))
    co.push_batch_with_filter([i32([1, 2, 3])], f)
    co.push_batch_with_filter([i32([4, 5, 6])], f)
    co.finish_buffered_batch()
    assert logical(co.next_completed_batch()[0]) == [1, 3, 4, 6]
    co.close()
    # coalesce.rs:271-288 push_batch_with_indices
    co = BatchCoalescer(gpu, [abi.I32], 1000)
    co.push_batch([i32([0, 0, 0])])
    co.push_batch_with_indices([i32([1, 1, 4, 5, 1, 4])], HostArray.from_list(abi.U64, [0, 1, 4, 2, 5, 3]))
    co.finish_buffered_batch()
    assert logical(co.next_completed_batch()[0]) == [0, 0, 0, 1, 1, 1, 4, 4, 5]
    co.close()
    # coalesce.rs:475-481 column-count check
    co = BatchCoalescer(gpu, [abi.I32, abi.I32], 8)
    with pytest.raises(acu.ArrowError) as e:
        co.push_batch([i32([1])])
    assert "Batch has 1 columns but BatchCoalescer expects 2" in str(e.value)
    co.close()


def test_single_batch_vs_target(gpu):
    """coalesce.rs:836-878: one large batch greater than / smaller than / equal to / a multiple of the target."""
    for rows, target, expect in [(4096, 1000, [1000, 1000, 1000, 1000, 96]), (4096, 8192, [4096]), (4096, 4096, [4096]), (4096, 1024, [1024] * 4)]:
        co = BatchCoalescer(gpu, [abi.U32], target)
        co.push_batch([HostArray.from_numpy(abi.U32, np.arange(rows, dtype=np.uint32))])
        co.finish_buffered_batch()
        got = drain(co)
        assert [b[0].length for b in got] == expect
        assert sum((logical(b[0]) for b in got), []) == list(range(rows))
        co.close()


@pytest.mark.parametrize("target", [7, 64, 1000, 4096])
def test_coalesce_fuzz(gpu, oracle, target):
    """Mixed schema, random batch sizes, plain / filtered / taken pushes: concatenation split at `target`."""
    rng = np.random.default_rng(1200 + target)
    schema = [abi.I64, abi.F64, BOOL, "utf8", abi.I8, abi.I32]
    co = BatchCoalescer(gpu, schema, target)
    expect_rows = [[] for _ in schema]

    def make(n):
        o, d, nl = rand_strings(rng, n, 0.15)
        return [rand_array(rng, abi.I64, n, 0.1), rand_array(rng, abi.F64, n, None, small=False), rand_bool(rng, n, 0.5, 0.2),
                Utf8Column(o, d, nl), rand_array(rng, abi.I8, n, 0.3, offset=3), rand_array(rng, abi.I32, n, 0.0)]

    for step in range(14):
        n = int(rng.choice([0, 1, 5, 63, 64, 65, 300, 1500, 5000]))
        cols = make(n)
        mode = step % 3
        if mode == 0:
            co.push_batch(cols)
            pushed = cols
        elif mode == 1:
            pred = rand_bool(rng, n, float(rng.choice([0.0, 0.05, 0.5, 1.0])), 0.05)
            co.push_batch_with_filter(cols, pred)
            pushed = [Utf8Column(*oracle.filter_bytes(c.offsets, c.data, c.nulls, pred)) if isinstance(c, Utf8Column) else oracle.filter(c, pred) for c in cols]
        else:
            m = int(rng.integers(0, 2 * n + 1)) if n else 0
            idx = HostArray.from_numpy(abi.U32, rng.integers(0, max(n, 1), m).astype(np.uint32), rng.random(m) >= 0.1)
            if n == 0:
                idx = HostArray.from_numpy(abi.U32, np.zeros(0, np.uint32))
            co.push_batch_with_indices(cols, idx)
            pushed = [Utf8Column(*oracle.take_bytes(c.offsets, c.data, c.nulls, idx)) if isinstance(c, Utf8Column) else oracle.take(c, idx) for c in cols]
        for k, c in enumerate(pushed):
            expect_rows[k] += logical(c)
        assert co.get_buffered_rows() < target
    co.finish_buffered_batch()
    got = drain(co)
    total = len(expect_rows[0])
    sizes = [b[0].length for b in got]
    assert sizes == [target] * (total // target) + ([total % target] if total % target else [])
    pos = 0
    for b in got:
        n = b[0].length
        for k in range(len(schema)):
            exp = expect_rows[k][pos:pos + n]
            g = logical(b[k])
            if schema[k] == abi.F64:
                assert all((x is None and y is None) or (x is not None and y is not None and (x == y or (x != x and y != y))) for x, y in zip(g, exp))
            else:
                assert g == exp, f"column {k} of the batch at row {pos}"
            # NullBufferBuilder: a NullBuffer only when a null was appended to THIS output batch
            assert has_nulls_buffer(b[k]) == any(x is None for x in exp), f"NullBuffer presence, column {k}"
        pos += n
    assert co.is_empty()
    co.close()


def test_bitmap_copy_and_fill_primitives(gpu):
    """acu_bitmap_copy / acu_bitmap_fill at every (source offset, destination offset, length) alignment class."""
    rng = np.random.default_rng(3)
    lib, h = gpu.lib, gpu.h
    nbits = 1024
    src_bits = rng.random(nbits) < 0.5
    d_src = gpu.malloc(nbits // 8 + 8)
    gpu.h2d(d_src, acu.pack_bits(src_bits))
    d_dst = gpu.malloc(nbits // 8 + 8)
    for soff, doff, ln in [(0, 0, 1), (3, 5, 60), (0, 64, 64), (7, 63, 2), (1, 0, 640), (13, 129, 517), (64, 1, 63), (5, 5, 1000), (0, 1023, 1)]:
        base = rng.random(nbits) < 0.5
        gpu.h2d(d_dst, acu.pack_bits(base))
        cnt = C.c_int64(0)
        gpu.check(lib.acu_bitmap_copy(h, d_src, soff, d_dst, doff, ln, C.byref(cnt)))
        got = acu.unpack_bits(gpu.d2h(d_dst, nbits // 8), 0, nbits)
        exp = base.copy()
        exp[doff:doff + ln] = src_bits[soff:soff + ln]
        assert np.array_equal(got, exp), (soff, doff, ln)
        assert cnt.value == int(src_bits[soff:soff + ln].sum())
        for value in (0, 1):
            gpu.h2d(d_dst, acu.pack_bits(base))
            gpu.check(lib.acu_bitmap_fill(h, d_dst, doff, ln, value))
            gpu.sync()
            got = acu.unpack_bits(gpu.d2h(d_dst, nbits // 8), 0, nbits)
            exp = base.copy()
            exp[doff:doff + ln] = bool(value)
            assert np.array_equal(got, exp), (doff, ln, value)
    gpu.free(d_src)
    gpu.free(d_dst)
