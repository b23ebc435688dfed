"""acu_filter_plan_slices = IterationStrategy::Slices of FilterBuilder::optimize / SlicesIterator (arrow-select/src/filter.rs:44-77,
:285-298) on the device vs the oracle (bit offsets, nulls, word / tile boundaries), the reference's cases (filter.rs:1640-1678)
and the sizing / capacity contract."""
import ctypes as C

import numpy as np
import pytest

from acu import _abi as abi
from acu import HostArray

pytestmark = pytest.mark.gpu


def test_reference_cases(gpu):
    assert gpu.filter_slices(HostArray.bool_from_numpy(np.array([i == 1 for i in range(64)]))) == [(1, 2)]
    assert gpu.filter_slices(HostArray.bool_from_numpy(np.array([i != 1 for i in range(64)]))) == [(0, 1), (2, 64)]
    assert gpu.filter_slices(HostArray.bool_from_numpy(np.array([i % 62 != 0 for i in range(130)]))) == [(1, 62), (63, 124), (125, 130)]


def test_fuzz(gpu, oracle):
    rng = np.random.default_rng(4)
    for n in [0, 1, 63, 64, 65, 127, 128, 1023, 1024, 1025, 4097, 70001, 300000]:
        for p in (0.0, 0.02, 0.5, 0.98, 1.0):
            bits = rng.random(n) < p
            if n > 200 and p > 0.9:
                bits[60:140] = True  # a run across several words
            mask = None if rng.random() < 0.3 else rng.random(n) >= 0.05
            f = HostArray.bool_from_numpy(bits, mask, bit_offset=int(rng.integers(0, 9)), mask_offset=int(rng.integers(0, 9)))
            assert gpu.filter_slices(f) == oracle.filter_slices(f), f"n={n} p={p}"


def test_capacity_contract(gpu):
    f = HostArray.bool_from_numpy(np.array([i % 3 != 0 for i in range(1000)]))
    dp = gpu.upload(f)
    plan = C.c_void_p()
    pd = dp.descriptor()
    gpu.check(gpu.lib.acu_filter_plan_create(gpu.h, C.byref(pd), C.byref(plan)))
    n = C.c_int64(0)
    gpu.check(gpu.lib.acu_filter_plan_slices(gpu.h, plan, None, 0, C.byref(n)))
    assert n.value == 333
    out = gpu.malloc(16 * 10)
    assert gpu.lib.acu_filter_plan_slices(gpu.h, plan, out, 10, C.byref(n)) == abi.ERR_INVALID_ARGUMENT
    gpu.free(out)
    gpu.lib.acu_filter_plan_destroy(gpu.h, plan)
    dp.free()
