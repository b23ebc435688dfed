"""concat / concat_batches on the device vs the oracle (arrow-select/src/concat.rs:495-640): values incl. the bytes under
null slots, validity, null_count, NullBuffer presence (NullBufferBuilder semantics), offsets, errors."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray, Utf8Column

from test_gpu_parity import assert_same, expect_same_error, rand_array, rand_bool
from test_oracle_cmp_bytes import rand_strings, utf8_column

pytestmark = pytest.mark.gpu


def assert_same_utf8(g, e, what):
    assert np.array_equal(g.offsets, e.offsets), f"{what}: offsets"
    assert np.array_equal(g.data[: int(e.offsets[-1]) if len(e.offsets) else 0], e.data[: int(e.offsets[-1]) if len(e.offsets) else 0]), f"{what}: bytes"
    assert (g.nulls.validity is None) == (e.nulls.validity is None), f"{what}: NullBuffer presence"
    assert g.nulls.null_count == e.nulls.null_count
    assert np.array_equal(g.nulls.valid_mask(), e.nulls.valid_mask())


@pytest.mark.parametrize("dtype", [abi.I8, abi.I32, abi.I64, abi.F64, BOOL])
def test_concat_fuzz(gpu, oracle, dtype):
    rng = np.random.default_rng(900 + (5 if dtype == BOOL else dtype))
    for trial in range(12):
        n_arrays = int(rng.integers(1, 6))
        arrays = []
        for _ in range(n_arrays):
            n = int(rng.choice([0, 1, 31, 64, 65, 1000, 5000]))
            null_p = [None, 0.0, 0.2][int(rng.integers(0, 3))]
            off = int(rng.integers(0, 4))
            arrays.append(rand_bool(rng, n, 0.5, null_p, off) if dtype == BOOL else rand_array(rng, dtype, n, null_p, off))
        assert_same(gpu.concat(arrays), oracle.concat(arrays), f"concat trial {trial} lens {[a.length for a in arrays]}")


@pytest.mark.parametrize("large", [False, True])
def test_concat_utf8_fuzz(gpu, oracle, large):
    rng = np.random.default_rng(950 + large)
    for trial in range(10):
        cols = []
        for _ in range(int(rng.integers(1, 5))):
            n = int(rng.choice([0, 1, 33, 500]))
            col = utf8_column(rand_strings(rng, n + 2, [None, 0.2][int(rng.integers(0, 2))]), large)
            if rng.random() < 0.5:  # a slice: offsets no longer start at 0
                col = Utf8Column(col.offsets[1: n + 2], col.data, col.nulls.slice(1, n))
                col.nulls.values = np.zeros(0, np.uint8)
            else:
                col = Utf8Column(col.offsets[: n + 1], col.data, col.nulls.slice(0, n))
                col.nulls.values = np.zeros(0, np.uint8)
            cols.append(col)
        assert_same_utf8(gpu.concat(cols), oracle.concat(cols), f"concat utf8 trial {trial}")


def test_concat_batches(gpu, oracle):
    rng = np.random.default_rng(7)
    batches = []
    for n in (100, 0, 1000, 65):
        s = utf8_column(rand_strings(rng, n, 0.1))
        batches.append([rand_array(rng, abi.I64, n, 0.1), rand_array(rng, abi.F64, n, None), rand_bool(rng, n, 0.5, 0.3), s])
    g, e = gpu.concat_batches(batches), oracle.concat_batches(batches)
    for c in range(3):
        assert_same(g[c], e[c], f"concat_batches column {c}")
    assert_same_utf8(g[3], e[3], "concat_batches utf8 column")


def test_concat_errors(gpu, oracle):
    got, _ = expect_same_error(gpu, oracle, lambda be: be.concat([]))
    assert got is None
    rng = np.random.default_rng(8)
    got, _ = expect_same_error(gpu, oracle, lambda be: be.concat([rand_array(rng, abi.I32, 4, None), rand_array(rng, abi.I64, 4, None)]))
    assert got is None
