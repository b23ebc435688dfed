"""Runs the transcribed reference test vectors (tests/golden/vectors.json) against a backend
(the CPU oracle or the GPU Context — both expose the same method names)."""
import json
import math
import os
import struct

import numpy as np

import acu
from acu import _abi as abi
from acu import BOOL, HostArray

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")
DT = {n: i for i, n in enumerate(abi.DTYPE_NAMES)}
DT["bool"] = BOOL


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def _f64_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def decode_value(x):
    if isinstance(x, str):
        return {"nan": math.nan, "+nan": _f64_bits(0x7FF8000000000000), "-nan": _f64_bits(0xFFF8000000000000),
                "inf": math.inf, "-inf": -math.inf, "-0.0": -0.0}[x]
    return x


def build_array(spec):
    dtype = DT[spec["dtype"]]
    if "raw_values" in spec:  # explicit values buffer + validity (values under nulls are preserved)
        vals = [decode_value(v) for v in spec["raw_values"]]
        mask = np.array(spec["raw_validity"], dtype=bool)
        if dtype == BOOL:
            h = HostArray.bool_from_numpy(np.array(vals, dtype=bool), mask)
        else:
            h = HostArray.from_numpy(dtype, np.array(vals, dtype=acu.NP_DTYPES[dtype]), mask)
    else:
        items = [None if v is None else decode_value(v) for v in spec["data"]]
        h = HostArray.from_list(dtype, items, force_validity=spec.get("force_validity", False))
        if dtype in (abi.F32, abi.F64):  # keep NaN sign bits exactly as requested
            for i, v in enumerate(spec["data"]):
                if v == "-nan":
                    h.values[i] = np.array([0xFFF8000000000000], dtype=np.uint64).view(np.float64)[0]
    if "slice" in spec:
        h = h.slice(spec["slice"][0], spec["slice"][1])
    if spec.get("scalar"):
        h = h.scalar()
    return h


def build_strings(spec):
    """{"strings": [...], "large": bool (i64 offsets), "slice": [offset, len] (value offsets then no longer start at 0)}"""
    strings = spec["strings"]
    offsets = np.zeros(len(strings) + 1, dtype=np.int64 if spec.get("large") else np.int32)
    chunks = []
    for i, s in enumerate(strings):
        b = b"" if s is None else s.encode()
        chunks.append(b)
        offsets[i + 1] = offsets[i] + len(b)
    data = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy()
    mask = [s is not None for s in strings]
    nulls = HostArray.from_list(abi.U8, [0 if m else None for m in mask])
    nulls.values = np.zeros(0, np.uint8)
    data = data[: offsets[-1] + 16]
    if "slice" in spec:  # Array::slice of a byte array: offsets window + validity bit offset, same data buffer
        off, ln = spec["slice"]
        offsets = offsets[off: off + ln + 1]
        nulls = nulls.slice(off, ln)
        nulls.values = np.zeros(0, np.uint8)
    return offsets, data, nulls


def _to_bytes(x):
    if x is None:
        return None
    return x.encode() if isinstance(x, str) else bytes(x)


def build_bytes_column(spec, large=False):
    """{"bytes": [str | [ints] | None ...], "scalar": bool, "slice": [offset, len]} -> acu.Utf8Column"""
    items = [_to_bytes(x) for x in spec["bytes"]]
    offsets = np.zeros(len(items) + 1, dtype=np.int64 if large else np.int32)
    for i, b in enumerate(items):
        offsets[i + 1] = offsets[i] + len(b or b"")
    data = np.frombuffer(b"".join(b or b"" for b in items) + b"\0" * 16, dtype=np.uint8).copy()
    nulls = HostArray.from_list(abi.U8, [0 if b is not None else None for b in items])
    nulls.values = np.zeros(0, np.uint8)
    if "slice" in spec:
        off, ln = spec["slice"]
        offsets = offsets[off: off + ln + 1]
        nulls = nulls.slice(off, ln)
        nulls.values = np.zeros(0, np.uint8)
    nulls.is_scalar = bool(spec.get("scalar"))
    return acu.Utf8Column(offsets, data, nulls)


def build_view_column(spec):
    return acu.ViewColumn.from_values([_to_bytes(x) for x in spec["bytes"]], 64, scalar=bool(spec.get("scalar")))


def strings_of(offsets, data, nulls):
    mask = nulls.valid_mask()
    out = []
    for i in range(nulls.length):
        out.append(bytes(data[offsets[i]: offsets[i + 1]]).decode() if mask[i] else None)
    return out


def _match(expected, got, f32=False):
    if expected == "*":
        return True
    if expected is None:
        return got is None
    if got is None:
        return False
    if isinstance(expected, str):
        if expected == "nan":
            return isinstance(got, float) and math.isnan(got)
        if expected in ("+nan", "-nan"):
            return isinstance(got, float) and math.isnan(got) and (math.copysign(1.0, got) < 0) == (expected == "-nan")
        expected = decode_value(expected)
    if isinstance(expected, bool) or isinstance(got, bool):
        return bool(expected) == bool(got)
    if isinstance(expected, float) or isinstance(got, float):
        e = float(np.float32(expected)) if f32 else float(expected)
        if e == 0.0 and float(got) == 0.0:
            return math.copysign(1.0, e) == math.copysign(1.0, float(got))
        return e == float(got)
    return int(expected) == int(got)


def check_array(expect, res):
    if "data" in expect:
        got = res.to_list()
        assert len(got) == len(expect["data"]), f"length {len(got)} != {len(expect['data'])}"
        for i, (e, g) in enumerate(zip(expect["data"], got)):
            assert _match(e, g, expect.get("f32", False)), f"slot {i}: expected {e!r}, got {g!r}"
    if "len" in expect:
        assert res.length == expect["len"]
    if "null_count" in expect:
        assert int((~res.valid_mask()).sum()) == expect["null_count"]
        assert (res.null_count if res.validity is not None else 0) == expect["null_count"]
    if "at" in expect:
        got = res.to_list()
        for k, e in expect["at"].items():
            assert _match(e, got[int(k)]), f"slot {k}: expected {e!r}, got {got[int(k)]!r}"
    if expect.get("always_validity"):
        assert res.validity is not None
    if expect.get("no_validity"):
        assert res.validity is None
    # cached null_count must agree with the bitmap (NullBuffer invariant)
    if res.validity is not None:
        assert res.null_count == int((~res.valid_mask()).sum())


def run_case(backend, c):
    op = c["op"]

    def call():
        if op == "filter":
            return backend.filter(build_array(c["values"]), build_array(c["predicate"]))
        if op == "selected":
            pred = build_array(c["predicate"])
            iota = HostArray.from_numpy(abi.I32, np.arange(pred.length, dtype=np.int32))
            return backend.filter(iota, pred), backend.filter_plan(pred)[0]
        if op == "filter_utf8":
            o, d, n = build_strings(c["values"])
            return backend.filter_bytes(o, d, n, build_array(c["predicate"]))
        if op == "take":
            return backend.take(build_array(c["values"]), build_array(c["indices"]), c.get("check_bounds", False))
        if op == "take_utf8":
            o, d, n = build_strings(c["values"])
            return backend.take_bytes(o, d, n, build_array(c["indices"]), c.get("check_bounds", False))
        if op in ("add", "add_wrapping", "sub", "sub_wrapping", "mul", "mul_wrapping", "div", "rem"):
            return getattr(backend, op)(build_array(c["a"]), build_array(c["b"]))
        if op in ("neg", "neg_wrapping"):
            return getattr(backend, op)(build_array(c["a"]))
        if op in ("eq", "neq", "lt", "lt_eq", "gt", "gt_eq", "distinct", "not_distinct"):
            return getattr(backend, op)(build_array(c["a"]), build_array(c["b"]))
        if op == "cast":
            return backend.cast(build_array(c["a"]), DT[c["to"]], c.get("safe", True))
        if op in ("sum", "min", "max", "sum_checked"):
            return getattr(backend, op)(build_array(c["a"]))
        if op in ("and_", "or_", "and_not", "and_kleene", "or_kleene"):
            return getattr(backend, op)(build_array(c["a"]), build_array(c["b"]))
        if op == "concat":
            return backend.concat([build_array(a) for a in c["arrays"]])
        if op == "concat_utf8":
            r = backend.concat([acu.Utf8Column(*build_strings(a)) for a in c["arrays"]])
            return r.offsets, r.data, r.nulls
        if op == "cmp_bytes":
            return backend.cmp_bytes(getattr(abi, c["cmp"].upper()), build_bytes_column(c["left"], c.get("large", False)),
                                     build_bytes_column(c["right"], c.get("large", False)))
        if op == "cmp_view":
            return backend.cmp_view(getattr(abi, c["cmp"].upper()), build_view_column(c["left"]), build_view_column(c["right"]))
        if op == "nullif":
            return backend.nullif(build_array(c["left"]), build_array(c["right"]))
        if op == "zip":
            return backend.zip(build_array(c["mask"]), build_array(c["truthy"]), build_array(c["falsy"]))
        if op in ("not_", "is_null", "is_not_null"):
            return getattr(backend, op)(build_array(c["a"]))
        raise AssertionError("unknown op " + op)

    if "expect_error" in c:
        try:
            call()
        except acu.ArrowError as e:
            assert str(e) == c["expect_error"], f"{str(e)!r} != {c['expect_error']!r}"
            return
        raise AssertionError("expected error: " + c["expect_error"])
    if c.get("expect_panic"):
        try:
            call()
        except acu.ArrowError as e:
            assert e.status == abi.ERR_PANIC_OUT_OF_BOUNDS
            return
        raise AssertionError("expected the reference's panic to surface as ACU_ERR_PANIC_OUT_OF_BOUNDS")
    res = call()
    e = c["expect"]
    if op == "selected":
        arr, count = res
        assert arr.to_list() == e["positions"]
        assert count == e["count"]
    elif op in ("filter_utf8", "take_utf8", "concat_utf8"):
        assert strings_of(*res) == e["strings"]
    elif op in ("sum", "min", "max", "sum_checked"):
        assert _match(e["scalar"], res), f"expected {e['scalar']!r}, got {res!r}"
    else:
        check_array(e, res)
