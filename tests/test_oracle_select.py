"""CPU: the oracle's nullif / zip restatements against the naive per-slot definitions the reference's own tests use
(nullif.rs:486-497 `test_nullif` helper: Some(true) => None, else the left slot; zip doc comment zip.rs:41-46)."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray

from test_gpu_parity import rand_array, rand_bool


@pytest.mark.parametrize("seed", range(6))
def test_nullif_matches_naive(oracle, seed):
    rng = np.random.default_rng(seed)
    for n in [0, 1, 63, 64, 65, 200, 1000]:
        left = rand_array(rng, abi.I32, n, [None, 0.2, 0.0][seed % 3], seed % 5)
        right = rand_bool(rng, n, 0.4, [None, 0.3][seed % 2], seed % 7)
        out = oracle.nullif(left, right)
        kill = right.value_array() & right.valid_mask()
        exp_valid = left.valid_mask() & ~kill
        assert np.array_equal(out.valid_mask(), exp_valid)
        assert np.array_equal(out.value_array(), left.value_array())  # values are shared untouched
        if n:  # an empty array is returned as it is (nullif.rs:54-56)
            assert (out.validity is None) == bool(exp_valid.all())
        if n and out.validity is not None:
            assert out.null_count == int((~exp_valid).sum())


@pytest.mark.parametrize("seed", range(6))
def test_zip_matches_naive(oracle, seed):
    rng = np.random.default_rng(100 + seed)
    for n in [0, 1, 63, 64, 65, 200, 1000]:
        t = rand_array(rng, abi.I64, n, [None, 0.2, 0.0][seed % 3], seed % 4)
        f = rand_array(rng, abi.I64, n, [0.1, None, None][seed % 3], seed % 3)
        m = rand_bool(rng, n, 0.5, [None, 0.3][seed % 2], seed % 5)
        out = oracle.zip(m, t, f)
        sel = m.value_array() & m.valid_mask()
        assert np.array_equal(out.value_array(), np.where(sel, t.value_array(), f.value_array()))  # bytes copied blindly
        exp_valid = np.where(sel, t.valid_mask(), f.valid_mask())
        assert np.array_equal(out.valid_mask(), exp_valid)
        has_nulls_in = (t.validity is not None and not t.valid_mask().all()) or (f.validity is not None and not f.valid_mask().all())
        assert (out.validity is not None) == bool(has_nulls_in and not exp_valid.all())


def test_zip_scalar_zipper_conventions(oracle):
    """PrimitiveScalarImpl (zip.rs:392-440): (Some, None) => every slot holds the truthy value and nulls = mask (always Some)."""
    m = HostArray.bool_from_numpy(np.array([True, False, True, True]))
    t = HostArray.from_list(abi.I32, [42]).scalar()
    fnull = HostArray.from_list(abi.I32, [None]).scalar()
    out = oracle.zip(m, t, fnull)
    assert out.value_array().tolist() == [42, 42, 42, 42] and out.to_list() == [42, None, 42, 42] and out.validity is not None
    out = oracle.zip(m, fnull, t)
    assert out.value_array().tolist() == [42, 42, 42, 42] and out.to_list() == [None, 42, None, None]
    allt = HostArray.bool_from_numpy(np.array([True, True]))
    out = oracle.zip(allt, t, fnull)
    assert out.validity is not None and out.null_count == 0  # the NullBuffer is kept even without nulls
    out = oracle.zip(m, fnull, fnull)
    assert out.value_array().tolist() == [0, 0, 0, 0] and out.null_count == 4
