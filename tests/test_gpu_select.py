"""nullif / zip / fused compare->filter on the device vs the oracle (bit-exact: values incl. the bytes under null slots,
validity bits, null_count, NullBuffer presence, errors). Reference: arrow-select/src/nullif.rs:44-113, zip.rs:99-440,
arrow-ord/src/cmp.rs:220-382 + arrow-select/src/filter.rs:254-273."""
import numpy as np
import pytest

import acu
from acu import _abi as abi
from acu import BOOL, HostArray

from test_gpu_parity import SIZES, assert_same, expect_same_error, rand_array, rand_bool

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [abi.I8, abi.I32, abi.I64, abi.F64, BOOL])
def test_nullif_fuzz(gpu, oracle, dtype):
    rng = np.random.default_rng(300 + (7 if dtype == BOOL else dtype))
    for n in SIZES:
        for null_p, r_null_p, off in [(None, None, 0), (0.1, 0.2, 3), (0.0, None, 1), (0.5, 0.5, 7), (None, 0.3, 2)]:
            left = rand_bool(rng, n, 0.5, null_p, off) if dtype == BOOL else rand_array(rng, dtype, n, null_p, off)
            for true_p in (0.0, 0.3, 1.0):
                right = rand_bool(rng, n, true_p, r_null_p, off + 1)
                assert_same(gpu.nullif(left, right), oracle.nullif(left, right), f"nullif n={n} nulls={null_p}/{r_null_p} p={true_p}")


def test_nullif_length_mismatch(gpu, oracle):
    rng = np.random.default_rng(1)
    got, _ = expect_same_error(gpu, oracle, lambda be: be.nullif(rand_array(rng, abi.I32, 10, 0.1), rand_bool(rng, 9, 0.5, None)))
    assert got is None


@pytest.mark.parametrize("dtype", [abi.I8, abi.I16, abi.I32, abi.I64, abi.F32, abi.F64])
def test_zip_arrays_fuzz(gpu, oracle, dtype):
    rng = np.random.default_rng(400 + dtype)
    for n in SIZES:
        for t_null, f_null, m_null, off in [(None, None, None, 0), (0.1, None, None, 1), (None, 0.2, 0.1, 2), (0.3, 0.3, 0.3, 5), (0.0, 0.0, None, 0)]:
            t, f = rand_array(rng, dtype, n, t_null, off), rand_array(rng, dtype, n, f_null, off and off - 1)
            for true_p in (0.0, 0.07, 0.5, 1.0):
                m = rand_bool(rng, n, true_p, m_null, off)
                assert_same(gpu.zip(m, t, f), oracle.zip(m, t, f), f"zip n={n} p={true_p} nulls={t_null}/{f_null}/{m_null}")


@pytest.mark.parametrize("dtype", [abi.I8, abi.I32, abi.I64, abi.F64])
def test_zip_scalars_fuzz(gpu, oracle, dtype):
    rng = np.random.default_rng(500 + dtype)
    for n in [0, 1, 63, 64, 65, 1000, 4097, 70001]:
        arr_ = rand_array(rng, dtype, n, 0.2, 3)
        for t_is_null in (False, True):
            for f_is_null in (False, True):
                ts = rand_array(rng, dtype, 1, 1.0 if t_is_null else None).scalar()
                fs = rand_array(rng, dtype, 1, 1.0 if f_is_null else None).scalar()
                for true_p, m_null in [(0.3, None), (0.5, 0.2), (1.0, None), (0.0, None)]:
                    m = rand_bool(rng, n, true_p, m_null, 2)
                    assert_same(gpu.zip(m, ts, fs), oracle.zip(m, ts, fs), f"zip scalar/scalar n={n} null={t_is_null}/{f_is_null}")
                    assert_same(gpu.zip(m, arr_, fs), oracle.zip(m, arr_, fs), f"zip array/scalar n={n}")
                    assert_same(gpu.zip(m, ts, arr_), oracle.zip(m, ts, arr_), f"zip scalar/array n={n}")


def test_zip_errors(gpu, oracle):
    rng = np.random.default_rng(2)
    m = rand_bool(rng, 10, 0.5, None)
    a10, a9 = rand_array(rng, abi.I32, 10, None), rand_array(rng, abi.I32, 9, None)
    for fn in (lambda be: be.zip(m, a9, a10), lambda be: be.zip(m, a10, a9)):
        got, _ = expect_same_error(gpu, oracle, fn)
        assert got is None


@pytest.mark.parametrize("dtype", [abi.I32, abi.I64, abi.F64, abi.U8])
@pytest.mark.parametrize("op", [abi.EQ, abi.NEQ, abi.LT, abi.LT_EQ, abi.GT, abi.GT_EQ, abi.DISTINCT, abi.NOT_DISTINCT])
def test_filter_cmp_fused_fuzz(gpu, oracle, dtype, op):
    """acu_filter_plan_create_cmp == acu_filter_plan_create(acu_cmp(...)): same selected rows, count and strategy."""
    rng = np.random.default_rng(600 + dtype * 10 + op)
    for n in [0, 1, 64, 65, 1000, 2047, 2048, 2049, 4096, 12345, 70001]:
        for a_null, b_null, off in [(None, None, 0), (0.1, 0.2, 3), (0.0, None, 0)]:
            a = rand_array(rng, dtype, n, a_null, off, small=True)
            b = rand_array(rng, dtype, n, b_null, off, small=True)
            values = rand_array(rng, abi.I64, n, 0.1, 1)
            (g, gplan), (e, eplan) = gpu.filter_cmp(values, op, a, b), oracle.filter_cmp(values, op, a, b)
            assert_same(g, e, f"filter_cmp n={n} op={op} nulls={a_null}/{b_null}")
            assert gplan == eplan
        if n:
            sc = rand_array(rng, dtype, 1, None, small=True).scalar()
            a = rand_array(rng, dtype, n, 0.1, 2, small=True)
            values = rand_array(rng, abi.F64, n, None)
            (g, gplan), (e, eplan) = gpu.filter_cmp(values, op, a, sc), oracle.filter_cmp(values, op, a, sc)
            assert_same(g, e, f"filter_cmp scalar n={n} op={op}")
            assert gplan == eplan
            nsc = rand_array(rng, dtype, 1, 1.0).scalar()  # null scalar: all-null predicate (or folded for distinct)
            (g, gplan), (e, eplan) = gpu.filter_cmp(values, op, nsc, a), oracle.filter_cmp(values, op, nsc, a)
            assert_same(g, e, f"filter_cmp null scalar n={n} op={op}")
            assert gplan == eplan


def test_filter_cmp_length_mismatch(gpu, oracle):
    rng = np.random.default_rng(3)
    a, b = rand_array(rng, abi.I32, 10, None), rand_array(rng, abi.I32, 11, None)
    got, _ = expect_same_error(gpu, oracle, lambda be: be.filter_cmp(a, abi.LT, a, b))
    assert got is None
