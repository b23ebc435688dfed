"""Arrow IPC stream -> HBM (acu_ipc_stream_*; reference arrow-ipc/src/reader.rs:1529-1671). The streams are written by an
independent Arrow implementation (pyarrow = Arrow C++, used here as the IPC *writer* peer, like tests/test_cdata.py uses it
as the C Data Interface peer); the decoded device columns are compared with the arrays that went in, buffer by buffer."""
import io

import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL

pa = pytest.importorskip("pyarrow")
pytestmark = pytest.mark.gpu


def stream_of(batches, schema):
    sink = io.BytesIO()
    with pa.ipc.new_stream(sink, schema) as w:
        for b in batches:
            w.write_batch(b)
    return sink.getvalue()


def check_column(col, arr):
    n = len(arr)
    valid = np.array([v is not None for v in arr.to_pylist()], dtype=bool)
    if pa.types.is_string(arr.type) or pa.types.is_binary(arr.type) or pa.types.is_large_string(arr.type):
        assert col.nulls.length == n
        got = []
        mask = col.nulls.valid_mask()
        for i in range(n):
            got.append(bytes(col.data[col.offsets[i]: col.offsets[i + 1]]) if mask[i] else None)
        exp = [None if v is None else (v.encode() if isinstance(v, str) else v) for v in arr.to_pylist()]
        assert got == exp
        assert (col.nulls.validity is None) == (arr.null_count == 0)
        return
    assert col.length == n
    assert np.array_equal(col.valid_mask(), valid)
    assert (col.validity is None) == (arr.null_count == 0)  # reader.rs:271: no NullBuffer when null_count == 0
    if arr.null_count:
        assert col.null_count == arr.null_count
    exp = arr.to_pylist()
    got = col.to_list()
    for g, e in zip(got, exp):
        assert (g is None) == (e is None)
        if e is not None:
            assert g == e or (isinstance(e, float) and np.isnan(e) and np.isnan(g))


def test_ipc_stream_roundtrip_types(gpu):
    rng = np.random.default_rng(0)
    n = 10_000
    def with_nulls(values, t, p=0.1):
        mask = rng.random(n) < p
        return pa.array([None if m else v for v, m in zip(values, mask)], type=t)
    cols = {
        "i8": with_nulls(rng.integers(-128, 127, n).tolist(), pa.int8()), "u16": with_nulls(rng.integers(0, 65535, n).tolist(), pa.uint16()),
        "i32": pa.array(rng.integers(-2**31, 2**31 - 1, n), type=pa.int32()), "i64": with_nulls(rng.integers(-2**62, 2**62, n).tolist(), pa.int64()),
        "u64": pa.array(rng.integers(0, 2**63, n).astype(np.uint64), type=pa.uint64()), "f32": with_nulls(rng.random(n).astype(np.float32).tolist(), pa.float32()),
        "f64": with_nulls(rng.random(n).tolist(), pa.float64(), 0.5), "b": with_nulls((rng.random(n) < 0.5).tolist(), pa.bool_()),
        "s": with_nulls(["s%d" % (i % 97) * (i % 5) for i in range(n)], pa.string()), "ls": with_nulls(["x" * (i % 7) for i in range(n)], pa.large_string()),
        "bin": with_nulls([bytes([i % 256]) * (i % 4) for i in range(n)], pa.binary()),
    }
    table = pa.table(cols)
    batches = table.to_batches(max_chunksize=3000)
    schema, got = gpu.ipc_read_stream(stream_of(batches, table.schema))
    assert [s[0] for s in schema] == list(cols)
    assert len(got) == len(batches)
    for gb, eb in zip(got, batches):
        for c, name in enumerate(cols):
            check_column(gb[c], eb.column(c))


def test_ipc_stream_feeds_the_hot_path_without_leaving_hbm(gpu, oracle):
    """decode -> filter_record_batch on the device views (no host round trip of the columns)."""
    import ctypes as C
    n = 50_000
    rng = np.random.default_rng(1)
    a = pa.array([None if m else int(v) for v, m in zip(rng.integers(-1000, 1000, n), rng.random(n) < 0.05)], type=pa.int64())
    f = pa.array(rng.random(n), type=pa.float64())
    table = pa.table({"a": a, "f": f})
    stream = stream_of(table.to_batches(max_chunksize=20_000), table.schema)
    totals = []

    def on_batch(cols, rows):  # predicate a > 0 built on the device, then filter_record_batch of both columns
        sc = acu.HostArray.from_list(abi.I64, [0]).scalar()
        dsc = gpu.upload(sc)
        out = gpu.alloc_out(acu.bitmap_bytes(rows), rows)
        scd = dsc.descriptor()
        gpu.check(gpu.lib.acu_cmp(gpu.h, abi.I64, abi.GT, C.byref(cols[0].array), C.byref(scd), C.byref(out)))
        pred = abi.Array()
        pred.values, pred.validity, pred.len = out.values, out.validity if out.has_validity else None, rows
        pred.null_count = out.null_count if out.has_validity else 0
        plan = C.c_void_p()
        gpu.check(gpu.lib.acu_filter_plan_create(gpu.h, C.byref(pred), C.byref(plan)))
        count = gpu.lib.acu_filter_plan_count(plan)
        outs = (abi.ColumnOut * 2)()
        for o in outs:
            o.array.values, o.array.validity = gpu.malloc(count * 8 + 16), gpu.malloc(acu.bitmap_bytes(count) + 8)
        gpu.check(gpu.lib.acu_filter_record_batch(gpu.h, plan, 2, cols, outs))
        vals = gpu.d2h(outs[0].array.values, count * 8, np.int64)
        gpu.lib.acu_filter_plan_destroy(gpu.h, plan)
        for o in outs:
            gpu.free(o.array.values)
            gpu.free(o.array.validity)
        gpu._free_out(out)
        dsc.free()
        return count, int(vals.sum())

    _, res = gpu.ipc_read_stream(stream, on_batch=on_batch)
    py = [v for v in a.to_pylist() if v is not None and v > 0]
    assert sum(r[0] for r in res) == len(py) and sum(r[1] for r in res) == sum(py)


def test_ipc_stream_errors(gpu):
    with pytest.raises(acu.ArrowError) as e:
        gpu.ipc_read_stream(b"")
    assert str(e.value) == "Ipc error: Expected schema message, found empty stream."
    t = pa.table({"a": pa.array([1, 2, 3])})
    s = stream_of(t.to_batches(), t.schema)
    # a stream that starts with the record batch message (schema stripped): reader.rs:1599-1604
    first_len = 8 + int(np.frombuffer(s[4:8], dtype=np.int32)[0])
    with pytest.raises(acu.ArrowError) as e:
        gpu.ipc_read_stream(s[first_len:])
    assert str(e.value) == "Ipc error: Expected a schema as the first message in the stream, got: RecordBatch"
    d = pa.table({"d": pa.array(["a", "b", "a"]).dictionary_encode()})
    with pytest.raises(acu.ArrowError) as e:
        gpu.ipc_read_stream(stream_of(d.to_batches(), d.schema))
    assert e.value.status == abi.ERR_NOT_YET_IMPLEMENTED and "dictionary" in str(e.value)
