"""cmp on Utf8 / LargeUtf8 / Binary and Utf8View / BinaryView operands on the device vs the oracle: value bits at every slot
(including under nulls), validity, null_count, NullBuffer presence, errors. Reference: arrow-ord/src/cmp.rs:220-382, :405-435,
:783-898. The oracle itself is pinned by the reference's comparison.rs vectors and against Python's bytes ordering
(tests/test_oracle_cmp_bytes.py)."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import ViewColumn

from test_gpu_parity import assert_same, expect_same_error
from test_oracle_cmp_bytes import rand_strings, utf8_column

pytestmark = pytest.mark.gpu

ALL_OPS = [abi.EQ, abi.NEQ, abi.LT, abi.LT_EQ, abi.GT, abi.GT_EQ, abi.DISTINCT, abi.NOT_DISTINCT]
SIZES = [0, 1, 31, 32, 33, 64, 127, 129, 1000, 4097, 20001]


@pytest.mark.parametrize("large", [False, True])
@pytest.mark.parametrize("op", ALL_OPS)
def test_cmp_bytes_fuzz(gpu, oracle, op, large):
    rng = np.random.default_rng(700 + op + 10 * large)
    for n in SIZES:
        for la_null, lb_null in [(None, None), (0.2, None), (0.1, 0.3)]:
            a, b = utf8_column(rand_strings(rng, n, la_null), large), utf8_column(rand_strings(rng, n, lb_null), large)
            assert_same(gpu.cmp_bytes(op, a, b), oracle.cmp_bytes(op, a, b), f"cmp_bytes n={n} op={op}")
        if n:
            a = utf8_column(rand_strings(rng, n, 0.1), large)
            for sc_items in ([rand_strings(rng, 1)[0]], [None]):
                sc = utf8_column(sc_items, large, scalar=True)
                assert_same(gpu.cmp_bytes(op, a, sc), oracle.cmp_bytes(op, a, sc), f"cmp_bytes array/scalar n={n} op={op}")
                assert_same(gpu.cmp_bytes(op, sc, a), oracle.cmp_bytes(op, sc, a), f"cmp_bytes scalar/array n={n} op={op}")


@pytest.mark.parametrize("op", ALL_OPS)
def test_cmp_view_fuzz(gpu, oracle, op):
    rng = np.random.default_rng(800 + op)
    garbage = rng.integers(0, 256, (7, 16), dtype=np.uint8)
    garbage[:, 1:4] = 0  # lengths < 256 so that a garbage view never points outside its buffers ...
    garbage[:, 0] = rng.integers(0, 13, 7)  # ... and stays inline
    for n in SIZES:
        for la_null, lb_null, short_only in [(None, None, False), (0.2, None, False), (0.1, 0.3, False), (None, 0.1, True)]:
            xa, xb = rand_strings(rng, n, la_null), rand_strings(rng, n, lb_null)
            if short_only:  # no data buffers on either side: the all-inline branches (cmp.rs:813-816, :867-871)
                xa = [None if x is None else x[:12] for x in xa]
                xb = [None if x is None else x[:12] for x in xb]
            a, b = ViewColumn.from_values(xa, 48, garbage_under_nulls=garbage), ViewColumn.from_values(xb, 96)
            assert_same(gpu.cmp_view(op, a, b), oracle.cmp_view(op, a, b), f"cmp_view n={n} op={op} short={short_only}")
        if n:
            a = ViewColumn.from_values(rand_strings(rng, n, 0.1), 64, garbage_under_nulls=garbage)
            for item in (b"", b"pre", b"pref", b"pref1", b"prefix-larger than 12 bytes string", None):
                sc = ViewColumn.from_values([item], scalar=True)
                assert_same(gpu.cmp_view(op, a, sc), oracle.cmp_view(op, a, sc), f"cmp_view array/scalar n={n} op={op} {item!r}")
                assert_same(gpu.cmp_view(op, sc, a), oracle.cmp_view(op, sc, a), f"cmp_view scalar/array n={n} op={op} {item!r}")


def test_cmp_bytes_length_mismatch(gpu, oracle):
    rng = np.random.default_rng(9)
    a, b = utf8_column(rand_strings(rng, 5)), utf8_column(rand_strings(rng, 6))
    got, _ = expect_same_error(gpu, oracle, lambda be: be.cmp_bytes(abi.EQ, a, b))
    assert got is None
    va, vb = ViewColumn.from_values(rand_strings(rng, 5)), ViewColumn.from_values(rand_strings(rng, 6))
    got, _ = expect_same_error(gpu, oracle, lambda be: be.cmp_view(abi.LT, va, vb))
    assert got is None
