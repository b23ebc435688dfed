"""Arrow C Data Interface / C Device Data Interface boundary (acu_export_column /
acu_import_column) checked against an independent implementation of the specification: pyarrow
(Arrow C++) imports what this library exports and exports what it imports. The structs are the
reference's FFI_ArrowArray / FFI_ArrowSchema (arrow-data/src/ffi.rs:37-69,
arrow-schema/src/ffi.rs); ArrowDeviceArray is the extension the reference lacks (SURVEY.md
§8(f) rank 4). No GPU needed: export / import move pointers and ownership, never bytes — here the
buffers are host memory with device_type = ARROW_DEVICE_CPU."""
import ctypes as C

import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray, Utf8Column

pa = pytest.importorskip("pyarrow")


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


def host_column(col):
    """acu_column over HOST buffers (numpy memory) of a HostArray / Utf8Column; returns (column, keepalive)."""
    c = abi.Column()
    keep = []
    if isinstance(col, Utf8Column):
        c.kind, c.width = abi.COL_BYTES, col.offsets.dtype.itemsize
        c.array = acu.host_descriptor(col.nulls)
        c.array.values = col.offsets.ctypes.data
        c.array.values_offset = 0
        c.data = col.data.ctypes.data
        keep += [col.offsets, col.data, col.nulls]
    else:
        c.kind = abi.COL_BOOLEAN if col.dtype == BOOL else abi.COL_PRIMITIVE
        c.width = 0 if col.dtype == BOOL else col.width()
        c.array = acu.host_descriptor(col)
        keep.append(col)
    return c, keep


def export(lib, col, dtype, released):
    colc, keep = host_column(col)
    arr, sch = abi.ArrowDeviceArray(), abi.ArrowSchema()
    cb = abi.RELEASE_OWNER(lambda owner: released.append(owner))
    st = lib.acu_export_column(None, C.byref(colc), dtype, abi.DEVICE_CPU, cb, 1234, C.byref(arr), C.byref(sch))
    assert st == abi.OK
    return arr, sch, (cb, keep)


@pytest.mark.parametrize("dtype,npdt", [(abi.I8, np.int8), (abi.I32, np.int32), (abi.I64, np.int64), (abi.U16, np.uint16), (abi.U64, np.uint64),
                                        (abi.F32, np.float32), (abi.F64, np.float64)])
def test_export_primitive_is_readable_by_pyarrow(lib, dtype, npdt):
    rng = np.random.default_rng(int(dtype))
    vals = rng.integers(0, 100, 50).astype(npdt)
    mask = rng.random(50) >= 0.3
    for h in (HostArray.from_numpy(dtype, vals, mask), HostArray.from_numpy(dtype, vals, None), HostArray.from_numpy(dtype, vals, mask, bit_offset=5).slice(7, 30)):
        released = []
        arr, sch, keep = export(lib, h, dtype, released)
        assert arr.device_type == abi.DEVICE_CPU and arr.array.length == h.length and arr.array.n_buffers == 2
        got = pa.Array._import_from_c(C.addressof(arr.array), C.addressof(sch))  # pyarrow takes ownership (moves the structs)
        assert got.to_pylist() == h.to_list()
        assert got.null_count == int((~h.valid_mask()).sum())
        assert released == []
        del got
        assert released == [1234]  # the consumer released exactly once -> our owner callback ran


def test_export_boolean_and_utf8(lib):
    rng = np.random.default_rng(5)
    b = HostArray.bool_from_numpy(rng.random(70) < 0.5, rng.random(70) >= 0.2, bit_offset=3, mask_offset=3).slice(9, 40)
    released = []
    arr, sch, keep = export(lib, b, abi.U8, released)
    assert sch.format == b"b" and arr.array.offset == 12
    got = pa.Array._import_from_c(C.addressof(arr.array), C.addressof(sch))
    assert got.to_pylist() == b.to_list()
    del got
    assert released == [1234]
    strings = ["", "a", None, "héllo", "x" * 40, None, "z"]
    offs = np.zeros(len(strings) + 1, dtype=np.int32)
    data = bytearray()
    for i, s in enumerate(strings):
        data += (s or "").encode()
        offs[i + 1] = len(data)
    nulls = HostArray(abi.U8, np.zeros(0, np.uint8), len(strings), acu.pack_bits(np.array([s is not None for s in strings])), 0, 0, 2)
    col = Utf8Column(offs, np.frombuffer(bytes(data) + b"\0" * 8, dtype=np.uint8).copy(), nulls)
    released = []
    arr, sch, keep = export(lib, col, abi.U8, released)
    assert sch.format == b"u" and arr.array.n_buffers == 3
    got = pa.Array._import_from_c(C.addressof(arr.array), C.addressof(sch))
    assert got.to_pylist() == strings and got.type == pa.utf8()
    del got
    assert released == [1234]


def test_export_as_device_array_roundtrips_through_pyarrow(lib):
    """The ArrowDeviceArray wrapper itself (device_type CPU here): pyarrow's C Device Data import."""
    h = HostArray.from_numpy(abi.I64, np.arange(20, dtype=np.int64), np.arange(20) % 3 != 0)
    released = []
    arr, sch, keep = export(lib, h, abi.I64, released)
    assert arr.device_id == -1 and not arr.sync_event
    got = pa.Array._import_from_c_device(C.addressof(arr), C.addressof(sch))
    assert got.to_pylist() == h.to_list()
    del got
    assert released == [1234]


def read_column(col, dtype, n):
    """Logical values of an imported host column, read back through its raw pointers."""
    a = col.array
    valid = np.ones(n, dtype=bool)
    if a.validity:
        nbytes = (a.validity_offset + n + 7) // 8
        bits = np.frombuffer((C.c_uint8 * nbytes).from_address(a.validity), dtype=np.uint8)
        valid = acu.unpack_bits(bits, a.validity_offset, n)
    if col.kind == abi.COL_BOOLEAN:
        nbytes = (a.values_offset + n + 7) // 8
        bits = np.frombuffer((C.c_uint8 * nbytes).from_address(a.values), dtype=np.uint8)
        vals = list(acu.unpack_bits(bits, a.values_offset, n))
    elif col.kind == abi.COL_BYTES:
        odt = np.int32 if col.width == 4 else np.int64
        offs = np.frombuffer((C.c_uint8 * ((n + 1) * col.width)).from_address(a.values), dtype=odt)
        data = np.frombuffer((C.c_uint8 * max(int(offs[-1]), 1)).from_address(col.data), dtype=np.uint8) if col.data else np.zeros(1, np.uint8)
        vals = [bytes(data[offs[i]:offs[i + 1]]).decode() for i in range(n)]
    else:
        npdt = acu.NP_DTYPES[dtype]
        vals = list(np.frombuffer((C.c_uint8 * (n * col.width)).from_address(a.values), dtype=npdt)) if n else []
    return ([(v.item() if hasattr(v, "item") else v) if ok else None for v, ok in zip(vals, valid)] if n else []), valid


@pytest.mark.parametrize("pa_arr,dtype", [
    (pa.array([1, None, 3, 4, None, 6, 7], type=pa.int32()), abi.I32),
    (pa.array([1.5, None, -0.0, float("inf")], type=pa.float64()), abi.F64),
    (pa.array([True, None, False, True, True, None, False, False, True], type=pa.bool_()), None),
    (pa.array(["a", None, "", "longer string", "ü"], type=pa.utf8()), None),
    (pa.array(["a", None, "bb"], type=pa.large_utf8()), None),
    (pa.array(list(range(100)), type=pa.uint8()), abi.U8),
])
def test_import_what_pyarrow_exports(lib, pa_arr, dtype):
    for view in (pa_arr, pa_arr.slice(1, len(pa_arr) - 2)):  # a slice exercises the single logical `offset`
        dev, sch = abi.ArrowDeviceArray(), abi.ArrowSchema()
        view._export_to_c_device(C.addressof(dev), C.addressof(sch))
        col, got_dtype = abi.Column(), C.c_int32(-1)
        assert lib.acu_import_column(None, C.byref(dev), C.byref(sch), C.byref(col), C.byref(got_dtype)) == abi.OK
        assert col.array.len == len(view)
        if dtype is not None:
            assert got_dtype.value == dtype
        vals, valid = read_column(col, got_dtype.value, len(view))
        exp = view.to_pylist()
        assert [v is not None for v in exp] == list(valid)
        for g, e in zip(vals, exp):
            assert (g is None and e is None) or g == e
        dev.array.release(C.byref(dev.array))  # consumer duty: release exactly once
        sch.release(C.byref(sch))
        assert not dev.array.release


def test_import_rejects_what_the_path_does_not_cover(lib):
    nested = pa.array([[1, 2], [3]], type=pa.list_(pa.int32()))
    dev, sch = abi.ArrowDeviceArray(), abi.ArrowSchema()
    nested._export_to_c_device(C.addressof(dev), C.addressof(sch))
    col, dt = abi.Column(), C.c_int32(0)
    assert lib.acu_import_column(None, C.byref(dev), C.byref(sch), C.byref(col), C.byref(dt)) == abi.ERR_NOT_YET_IMPLEMENTED
    dev.array.release(C.byref(dev.array))
    sch.release(C.byref(sch))


@pytest.mark.gpu
def test_device_array_export_import_and_compute(gpu, oracle):
    """A device-resident filter result leaves as ArrowDeviceArray(CUDA), comes back through acu_import_column and is
    consumed by the next kernel without touching the host; pyarrow's filter is the labelled secondary cross-check."""
    from test_gpu_parity import assert_same, rand_array, rand_bool
    rng = np.random.default_rng(99)
    n = 20000
    vals, pred = rand_array(rng, abi.I64, n, 0.1), rand_bool(rng, n, 0.4, None)
    dv, dp = gpu.upload(vals), gpu.upload(pred)
    plan = C.c_void_p()
    pd, vd = dp.descriptor(), dv.descriptor()
    gpu.check(gpu.lib.acu_filter_plan_create(gpu.h, C.byref(pd), C.byref(plan)))
    count = gpu.lib.acu_filter_plan_count(plan)
    out = gpu.alloc_out(count * 8, count)
    gpu.check(gpu.lib.acu_filter_primitive(gpu.h, plan, 8, C.byref(vd), C.byref(out)))
    gpu.lib.acu_filter_plan_destroy(gpu.h, plan)
    col = abi.Column()
    col.kind, col.width = abi.COL_PRIMITIVE, 8
    col.array.values, col.array.validity = out.values, out.validity if out.has_validity else None
    col.array.len, col.array.null_count = out.len, out.null_count if out.has_validity else 0
    released = []
    cb = abi.RELEASE_OWNER(lambda owner: released.append(owner))
    dev, sch = abi.ArrowDeviceArray(), abi.ArrowSchema()
    gpu.check(gpu.lib.acu_export_column(gpu.h, C.byref(col), abi.I64, abi.DEVICE_CUDA, cb, 7, C.byref(dev), C.byref(sch)))
    assert dev.device_type == abi.DEVICE_CUDA and dev.device_id == 0 and sch.format == b"l" and dev.array.length == count
    back, dt = abi.Column(), C.c_int32(-1)
    assert gpu.lib.acu_import_column(gpu.h, C.byref(dev), C.byref(sch), C.byref(back), C.byref(dt)) == abi.OK
    assert dt.value == abi.I64 and back.array.values == out.values  # zero copy: the very same device pointer
    # C Device Data Interface rules: device memory needs a ctx, and the right device
    assert gpu.lib.acu_import_column(None, C.byref(dev), C.byref(sch), C.byref(abi.Column()), C.byref(C.c_int32())) == abi.ERR_INVALID_ARGUMENT
    dev.device_id = 5
    with pytest.raises(acu.ArrowError) as e:
        gpu.check(gpu.lib.acu_import_column(gpu.h, C.byref(dev), C.byref(sch), C.byref(abi.Column()), C.byref(C.c_int32())))
    assert "CUDA device 5" in str(e.value)
    dev.device_id = 0
    try:  # a producer-side event: the import must order the ctx stream behind it
        from cuda.bindings import runtime as cudart
        err, ev = cudart.cudaEventCreate()
        assert int(err) == 0
        cudart.cudaEventRecord(ev, 0)
        holder = C.c_void_p(int(ev))
        dev.sync_event = C.addressof(holder)
        assert gpu.lib.acu_import_column(gpu.h, C.byref(dev), C.byref(sch), C.byref(back), C.byref(dt)) == abi.OK
        dev.sync_event = None
        cudart.cudaEventDestroy(ev)
    except ImportError:
        pass
    # next kernel on the imported column: sum
    bits, cnt = C.c_uint64(0), C.c_int64(0)
    gpu.check(gpu.lib.acu_aggregate(gpu.h, abi.I64, abi.SUM, C.byref(back.array), C.byref(bits), C.byref(cnt)))
    exp = oracle.filter(vals, pred)
    assert np.array([bits.value], dtype=np.uint64).view(np.int64)[0] == (oracle.sum(exp) or 0)
    assert_same(gpu.download_out(out, abi.I64), exp, "device result")  # frees the buffers
    dev.array.release(C.byref(dev.array))
    sch.release(C.byref(sch))
    assert released == [7]
    import pyarrow.compute as pc  # secondary cross-check (Arrow C++, not the oracle): logical values only
    pav = pa.array(vals.to_list(), type=pa.int64())
    assert pc.filter(pav, pa.array(pred.to_list()), null_selection_behavior="drop").to_pylist() == exp.to_list()
    dv.free()
    dp.free()


def test_export_of_a_device_column_needs_a_ctx_and_leaves_nothing_to_release(lib):
    """A failed export must not publish a live-looking ArrowArray (release stays NULL), and a device column cannot be
    exported without the ctx whose stream has to be synchronised first."""
    vals = np.arange(4, dtype=np.int64)
    colc = abi.Column()
    colc.kind, colc.width = abi.COL_PRIMITIVE, 8
    colc.array.values, colc.array.len = vals.ctypes.data, 4
    arr, sch = abi.ArrowDeviceArray(), abi.ArrowSchema()
    cb = abi.RELEASE_OWNER(lambda owner: None)
    assert lib.acu_export_column(None, C.byref(colc), abi.I64, abi.DEVICE_CUDA, cb, 1, C.byref(arr), C.byref(sch)) == abi.ERR_INVALID_ARGUMENT
    assert not arr.array.release
