"""CPU: the oracle's byte / byte-view comparisons (arrow-ord/src/cmp.rs:783-898, :405-435) against Python's own `bytes`
ordering (lexicographic on unsigned bytes, then length — the definition of Rust's `&[u8]` Ord) on random strings, with
shared prefixes, inline (<= 12 B) and out-of-line views, several data buffers, nulls and scalars."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import HostArray, Utf8Column, ViewColumn

OPS = [("eq", abi.EQ, lambda a, b: a == b), ("neq", abi.NEQ, lambda a, b: a != b), ("lt", abi.LT, lambda a, b: a < b),
       ("lt_eq", abi.LT_EQ, lambda a, b: a <= b), ("gt", abi.GT, lambda a, b: a > b), ("gt_eq", abi.GT_EQ, lambda a, b: a >= b)]


def rand_strings(rng, n, null_p=None):
    stems = [b"", b"a", b"pref", b"prefix-larger than 12", b"prefix-larger but", b"\xff\xf8", b"\x00", b"zz"]
    out = []
    for _ in range(n):
        if null_p is not None and rng.random() < null_p:
            out.append(None)
            continue
        s = stems[rng.integers(0, len(stems))] + bytes(rng.integers(0, 256, int(rng.integers(0, 20)), dtype=np.uint8) if rng.random() < 0.6 else [])
        out.append(s[: int(rng.integers(0, len(s) + 1))] if rng.random() < 0.3 else s)
    return out


def utf8_column(items, large=False, scalar=False):
    offs = np.zeros(len(items) + 1, dtype=np.int64 if large else np.int32)
    chunks = []
    for i, it in enumerate(items):
        b = b"" if it is None else it
        chunks.append(b)
        offs[i + 1] = offs[i] + len(b)
    data = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8).copy()
    nulls = HostArray.from_list(abi.U8, [0 if it is not None else None for it in items])
    nulls.values = np.zeros(0, np.uint8)
    nulls.is_scalar = scalar
    return Utf8Column(offs, data, nulls)


def expect(pyop, la, lb, ls, rs):
    n = len(lb) if ls else len(la)
    vals, valid = [], []
    for i in range(n):
        a, b = la[0 if ls else i], lb[0 if rs else i]
        valid.append(a is not None and b is not None)
        vals.append(pyop(a or b"", b or b"") if (a is not None and b is not None) else None)
    return vals, valid


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("kind", ["utf8", "large", "view"])
def test_cmp_bytes_matches_python_ordering(oracle, seed, kind):
    rng = np.random.default_rng(seed)
    for n in [1, 5, 63, 64, 65, 200]:
        for null_p in (None, 0.2):
            la, lb = rand_strings(rng, n, null_p), rand_strings(rng, n, null_p)
            sc = [rand_strings(rng, 1)[0]]
            for name, op, pyop in OPS:
                for ls, rs in [(False, False), (False, True), (True, False)]:
                    xa, xb = (sc if ls else la), (sc if rs else lb)
                    if kind == "view":
                        A, B = ViewColumn.from_values(xa, 40, scalar=ls), ViewColumn.from_values(xb, 64, scalar=rs)
                        res = oracle.cmp_view(op, A, B)
                    else:
                        A, B = utf8_column(xa, kind == "large", ls), utf8_column(xb, kind == "large", rs)
                        res = oracle.cmp_bytes(op, A, B)
                    vals, valid = expect(pyop, xa, xb, ls, rs)
                    got = res.to_list()
                    assert len(got) == len(vals)
                    for i, (g, e, v) in enumerate(zip(got, vals, valid)):
                        assert (g is None) == (not v), f"{name} {kind} slot {i} validity"
                        if v:
                            assert g == e, f"{name} {kind} slot {i}: {xa[0 if ls else i]!r} vs {xb[0 if rs else i]!r}"


def test_view_short_constant_fast_path_and_inline_only_arrays(oracle):
    """eq_inline_scalar (cmp.rs:405-435) and the all-inline branches (cmp.rs:813-816, :867-871)."""
    arr = ViewColumn.from_values([b"pref", b"pre", b"pref1", b"", None, b"prefix-larger than 12 bytes string", b"pref"], 64)
    for needle, exp in [(b"pref", [True, False, False, False, None, False, True]), (b"", [False, False, False, True, None, False, False])]:
        sc = ViewColumn.from_values([needle], scalar=True)
        assert oracle.cmp_view(abi.EQ, arr, sc).to_list() == exp
        assert oracle.cmp_view(abi.NEQ, sc, arr).to_list() == [None if e is None else (not e) for e in exp]
    a = ViewColumn.from_values([b"abc", b"abd", b"ab", b"", b"abcdefghijkl"])
    b = ViewColumn.from_values([b"abc", b"abc", b"abc", b"a", b"abcdefghijk"])
    assert a.buffers == [] and b.buffers == []
    assert oracle.cmp_view(abi.EQ, a, b).to_list() == [True, False, False, False, False]
    assert oracle.cmp_view(abi.LT, a, b).to_list() == [False, False, True, True, False]
    assert oracle.cmp_view(abi.GT_EQ, a, b).to_list() == [True, True, False, False, True]
