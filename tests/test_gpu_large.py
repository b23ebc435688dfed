"""Maximum-size / 64-bit-indexing checks at sizes the oracle cannot reach in seconds
(> 2^32 rows): verified through size-independent properties — the splitmix64 generator is
reproducible on the host for any row, so selected rows and tails are
spot-checked exactly, and counts must agree with popcounts."""
import ctypes as C

import numpy as np
import pytest

import acu
from acu import _abi as abi

pytestmark = pytest.mark.gpu

MASK64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def gen_bytes(seed, byte_index):
    """byte i of a kind-0 (raw u64) generated buffer."""
    with np.errstate(over="ignore"):
        words = splitmix64(np.uint64(seed) ^ (byte_index // 8).astype(np.uint64))
    return ((words >> ((byte_index % 8) * 8).astype(np.uint64)) & np.uint64(0xFF)).astype(np.uint8)


def make_arr(values, validity, n, nc, voff=0):
    a = abi.Array()
    a.values, a.values_offset, a.validity, a.validity_offset, a.len, a.null_count, a.is_scalar = values, voff, validity, 0, n, nc, 0
    return a


def test_filter_int8_beyond_2_pow_32_rows(gpu):
    n = (1 << 32) + 12_345_679  # rows of Int8: row indices need 33 bits
    lib, h = gpu.lib, gpu.h
    d_vals = gpu.malloc(n + 64)
    gpu.check(lib.acu_generate_values(h, 0, 42, 0, 0, d_vals, (n + 7) // 8))
    d_pred = gpu.malloc(abi.bitmap_bytes(n))
    gpu.check(lib.acu_generate_bits(h, 46, 0, 0.001, d_pred, n))
    pred = make_arr(d_pred, None, n, 0)
    plan = C.c_void_p()
    gpu.check(lib.acu_filter_plan_create(h, C.byref(pred), C.byref(plan)))
    count = lib.acu_filter_plan_count(plan)
    popc = C.c_int64(0)
    gpu.check(lib.acu_bitmap_count(h, d_pred, 0, None, 0, n, C.byref(popc)))
    assert count == popc.value and abs(count - n * 0.001) < 5 * (n * 0.001) ** 0.5 + 10
    d_idx = gpu.malloc(count * 8)
    gpu.check(lib.acu_filter_plan_indices(h, plan, abi.U64, d_idx))
    out = gpu.alloc_out(count, count)
    vals = make_arr(d_vals, None, n, 0)
    gpu.check(lib.acu_filter_primitive(h, plan, 1, C.byref(vals), C.byref(out)))
    idx = gpu.d2h(d_idx, count * 8, np.uint64)
    got = gpu.d2h(out.values, count, np.uint8)
    assert out.len == count and out.has_validity == 0
    assert np.all(np.diff(idx.astype(np.int64)) > 0) and int(idx[-1]) < n and int(idx[-1]) > (1 << 32)  # ascending, reaches past 2^32
    assert np.array_equal(got, gen_bytes(42, idx))  # out[k] == values[idx[k]] for every selected row
    # the indices really are the set bits: re-derive the predicate bit of each selected row on the host
    with np.errstate(over="ignore"):
        bit = splitmix64(np.uint64(46) ^ idx) < np.uint64(int(0.001 * 18446744073709551616.0))
    assert bit.all()
    # take with those u64 indices must reproduce the filter output (filter == take of the selected rows)
    out2 = gpu.alloc_out(count, count)
    ix = make_arr(d_idx, None, count, 0)
    gpu.check(lib.acu_take_primitive(h, 1, C.byref(vals), C.byref(ix), abi.U64, 1, C.byref(out2)))
    assert np.array_equal(gpu.d2h(out2.values, count, np.uint8), got)
    lib.acu_filter_plan_destroy(h, plan)
    for p in (d_vals, d_pred, d_idx):
        gpu.free(p)
    gpu._free_out(out)
    gpu._free_out(out2)


def test_add_int8_beyond_2_pow_32_rows(gpu):
    n = (1 << 32) + 1_000_003
    lib, h = gpu.lib, gpu.h
    d_a, d_b, d_av = gpu.malloc(n + 64), gpu.malloc(n + 64), gpu.malloc(abi.bitmap_bytes(n))
    gpu.check(lib.acu_generate_values(h, 0, 1, 0, 0, d_a, (n + 7) // 8))
    gpu.check(lib.acu_generate_values(h, 0, 2, 0, 0, d_b, (n + 7) // 8))
    gpu.check(lib.acu_generate_bits(h, 3, 0, 0.95, d_av, n))
    out = gpu.alloc_out(n, n)
    a, b = make_arr(d_a, d_av, n, -1), make_arr(d_b, None, n, 0)
    gpu.check(lib.acu_arith(h, abi.I8, abi.ADD_WRAPPING, C.byref(a), C.byref(b), C.byref(out)))
    valid = C.c_int64(0)
    gpu.check(lib.acu_bitmap_count(h, d_av, 0, None, 0, n, C.byref(valid)))
    assert out.len == n and out.has_validity == 1 and out.null_count == n - valid.value
    ovalid = C.c_int64(0)
    gpu.check(lib.acu_bitmap_count(h, out.validity, 0, None, 0, n, C.byref(ovalid)))
    assert ovalid.value == valid.value  # the output bitmap is the input bitmap (union with a no-null side)
    # spot-check three windows: start, around 2^32, and the ragged tail
    for lo in (0, (1 << 32) - 500, n - 777):
        m = min(1000, n - lo)
        idx = np.arange(lo, lo + m, dtype=np.uint64)
        exp = (gen_bytes(1, idx).astype(np.int16) + gen_bytes(2, idx).astype(np.int16)).astype(np.uint8)
        got = np.empty(m, dtype=np.uint8)
        gpu.check(lib.acu_memcpy_d2h(h, got.ctypes.data, out.values + lo, m))
        assert np.array_equal(got, exp), f"window at {lo}"
    for p in (d_a, d_b, d_av):
        gpu.free(p)
    gpu._free_out(out)
