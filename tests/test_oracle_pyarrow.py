"""SECONDARY cross-check of the oracle (labelled as such in SURVEY.md §8(c)): pyarrow is Arrow C++,
a different implementation whose semantics differ from arrow-rs in places (NaN / signed-zero
comparison, range-checked int->float casts, bytes under nulls), so it is NOT the oracle. Where the
two implementations are documented to agree — the LOGICAL result (value or null per slot) of
filter (nulls dropped), take, wrapping arithmetic on in-range data, comparisons away from NaN and
±0, casts within ±2^53, Kleene logic, sum / min / max on in-range integers — this test checks the
CPU oracle against it on seeded random inputs. It guards against a restatement error that the
233 transcribed reference vectors might not reach; the bit-level conventions are pinned by those
vectors, not here."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray

pa = pytest.importorskip("pyarrow")
pc = pytest.importorskip("pyarrow.compute")

PA_TYPES = {abi.I8: pa.int8(), abi.I16: pa.int16(), abi.I32: pa.int32(), abi.I64: pa.int64(), abi.U8: pa.uint8(), abi.U16: pa.uint16(),
            abi.U32: pa.uint32(), abi.U64: pa.uint64(), abi.F32: pa.float32(), abi.F64: pa.float64()}


def to_pa(h):
    return pa.array(h.to_list(), type=pa.bool_() if h.dtype == BOOL else PA_TYPES[h.dtype])


def rand_host(rng, dtype, n, null_p, lo=-1000, hi=1000):
    npdt = acu.NP_DTYPES[dtype]
    if dtype in (abi.F32, abi.F64):
        vals = (rng.integers(lo, hi, n) + rng.integers(0, 4, n) / 4.0).astype(npdt)  # exactly representable, no NaN, no -0.0
        vals[vals == 0] = 1.0
    else:
        info = np.iinfo(npdt)
        vals = rng.integers(max(lo, info.min), min(hi, info.max), n).astype(npdt)
    mask = None if null_p is None else rng.random(n) >= null_p
    h = HostArray.from_numpy(dtype, vals, mask, bit_offset=int(rng.integers(0, 9)) if mask is not None else 0)
    return h


def rand_bool_host(rng, n, p, null_p):
    return HostArray.bool_from_numpy(rng.random(n) < p, None if null_p is None else rng.random(n) >= null_p, bit_offset=int(rng.integers(0, 9)),
                                     mask_offset=int(rng.integers(0, 9)))


@pytest.mark.parametrize("dtype", [abi.I8, abi.I32, abi.I64, abi.U16, abi.F64])
def test_filter_and_take_logical(oracle, dtype):
    rng = np.random.default_rng(100 + dtype)
    for n in [0, 1, 65, 1000, 5000]:
        v = rand_host(rng, dtype, n, 0.2)
        pred = rand_bool_host(rng, n, 0.4, 0.1)
        got = oracle.filter(v, pred).to_list()
        exp = pc.filter(to_pa(v), to_pa(pred), null_selection_behavior="drop").to_pylist()  # arrow-rs: a null predicate slot selects nothing
        assert got == exp
        if n:
            idx = HostArray.from_numpy(abi.U32, rng.integers(0, n, 700).astype(np.uint32), rng.random(700) >= 0.15)
            assert oracle.take(v, idx).to_list() == pc.take(to_pa(v), to_pa(idx)).to_pylist()
            b = rand_bool_host(rng, n, 0.5, 0.2)
            assert oracle.take(b, idx).to_list() == pc.take(to_pa(b), to_pa(idx)).to_pylist()
            assert oracle.filter(b, pred).to_list() == pc.filter(to_pa(b), to_pa(pred), null_selection_behavior="drop").to_pylist()


def test_strings_filter_take_and_dictionary_decode(oracle):
    from golden_util import strings_of
    rng = np.random.default_rng(7)
    words = ["".join(chr(c) for c in rng.integers(97, 123, rng.integers(0, 15))) for _ in range(400)]
    strings = [w if rng.random() > 0.15 else None for w in words]
    offs = np.zeros(len(strings) + 1, dtype=np.int32)
    buf = bytearray()
    for i, s in enumerate(strings):
        buf += (s or "").encode()
        offs[i + 1] = len(buf)
    data = np.frombuffer(bytes(buf) + b"\0" * 16, dtype=np.uint8).copy()
    mask = np.array([s is not None for s in strings])
    nulls = HostArray(abi.U8, np.zeros(0, np.uint8), len(strings), acu.pack_bits(mask), 0, 0, int((~mask).sum()))
    pas = pa.array(strings, type=pa.utf8())
    pred = rand_bool_host(rng, len(strings), 0.5, 0.1)
    assert strings_of(*oracle.filter_bytes(offs, data, nulls, pred)) == pc.filter(pas, to_pa(pred), null_selection_behavior="drop").to_pylist()
    keys = HostArray.from_numpy(abi.I32, rng.integers(0, len(strings), 3000).astype(np.int32), rng.random(3000) >= 0.1)
    taken = strings_of(*oracle.take_bytes(offs, data, nulls, keys))
    assert taken == pc.take(pas, to_pa(keys)).to_pylist()
    # Dictionary<Int32, Utf8> -> Utf8 is exactly that take (arrow-cast/src/cast/dictionary.rs:310-317)
    dict_arr = pa.DictionaryArray.from_arrays(to_pa(keys), pas)
    assert taken == dict_arr.cast(pa.utf8()).to_pylist()


@pytest.mark.parametrize("dtype", [abi.I32, abi.I64, abi.U32, abi.F64])
def test_arithmetic_and_comparison_logical(oracle, dtype):
    rng = np.random.default_rng(200 + dtype)
    for n in [0, 1, 64, 999, 4097]:
        a, b = rand_host(rng, dtype, n, 0.2, 1, 1000), rand_host(rng, dtype, n, 0.2, 1, 1000)
        pa_a, pa_b = to_pa(a), to_pa(b)
        for op, pcf in (("add", pc.add), ("sub", pc.subtract), ("mul", pc.multiply), ("add_wrapping", pc.add), ("mul_wrapping", pc.multiply)):
            if op.startswith("sub") and dtype == abi.U32:
                continue  # unsigned underflow: checked in arrow-rs, wrapping in pyarrow's unchecked kernel
            assert getattr(oracle, op)(a, b).to_list() == pcf(pa_a, pa_b).to_pylist(), op
        if dtype in (abi.I32, abi.I64):
            assert oracle.div(a, b).to_list() == pc.divide(pa_a, pa_b).to_pylist()  # both truncate toward zero; b >= 1
        for op, pcf in (("eq", pc.equal), ("neq", pc.not_equal), ("lt", pc.less), ("lt_eq", pc.less_equal), ("gt", pc.greater), ("gt_eq", pc.greater_equal)):
            assert getattr(oracle, op)(a, b).to_list() == pcf(pa_a, pa_b).to_pylist(), op
        s = HostArray.from_list(dtype, [7], scalar=True)
        assert oracle.add(a, s).to_list() == pc.add(pa_a, pa.scalar(7, PA_TYPES[dtype])).to_pylist()
        assert oracle.lt(a, s).to_list() == pc.less(pa_a, pa.scalar(7, PA_TYPES[dtype])).to_pylist()


def test_cast_and_aggregates_logical(oracle):
    rng = np.random.default_rng(300)
    for n in [0, 1, 63, 1000]:
        a = rand_host(rng, abi.I64, n, 0.2, -2**52, 2**52)
        assert oracle.cast(a, abi.F64).to_list() == to_pa(a).cast(pa.float64()).to_pylist()
        small = rand_host(rng, abi.I64, n, 0.2, -100, 100)
        assert oracle.cast(small, abi.I8).to_list() == to_pa(small).cast(pa.int8()).to_pylist()
        f = rand_host(rng, abi.F64, n, 0.2, -100, 100)
        assert oracle.cast(f, abi.I32).to_list() == pc.cast(to_pa(f), pa.int32(), safe=False).to_pylist()  # truncation toward zero
        for col in (a, small):
            pav = to_pa(col)
            assert oracle.sum(col) == pc.sum(pav).as_py()
            assert oracle.min(col) == pc.min(pav).as_py() and oracle.max(col) == pc.max(pav).as_py()
        assert oracle.min(f) == pc.min(to_pa(f)).as_py() and oracle.max(f) == pc.max(to_pa(f)).as_py()


def test_boolean_kernels_logical(oracle):
    rng = np.random.default_rng(400)
    for n in [0, 1, 64, 1000]:
        for an, bn in [(None, None), (0.3, None), (None, 0.3), (0.3, 0.3)]:
            a, b = rand_bool_host(rng, n, 0.5, an), rand_bool_host(rng, n, 0.5, bn)
            pa_a, pa_b = to_pa(a), to_pa(b)
            assert oracle.and_kleene(a, b).to_list() == pc.and_kleene(pa_a, pa_b).to_pylist()
            assert oracle.or_kleene(a, b).to_list() == pc.or_kleene(pa_a, pa_b).to_pylist()
            assert oracle.and_(a, b).to_list() == pc.and_(pa_a, pa_b).to_pylist()
            assert oracle.or_(a, b).to_list() == pc.or_(pa_a, pa_b).to_pylist()
            assert oracle.and_not(a, b).to_list() == pc.and_not(pa_a, pa_b).to_pylist()
            assert oracle.not_(a).to_list() == pc.invert(pa_a).to_pylist()
            assert oracle.is_null(a).to_list() == pc.is_null(pa_a).to_pylist()
            assert oracle.is_not_null(a).to_list() == pc.is_valid(pa_a).to_pylist()
