"""BASELINE.json configs[0]: `arrow::compute::filter` on Int32Array len = 1e6, 50 % selectivity, CPU reference
(plumbing / correctness, no GPU). The oracle (the CPU restatement of arrow-select/src/filter.rs) runs the config and is
checked against the naive definition of filter — the reference's own cross-check, `filter_rust`
(arrow-select/src/filter.rs:1844-1886: zip values with the predicate, keep where Some(true)) — restated in numpy."""
import numpy as np

import acu
from acu import _abi as abi
from acu import BOOL, HostArray

N = 1_000_000


def config1(oracle, null_p, pred_null_p):
    vals = oracle.generate_values(4, 42, 0, 2**31 - 1, N, np.int32)
    validity = oracle.generate_bits(44, 0, 1 - null_p, N) if null_p is not None else None
    nc = -1 if validity is not None else 0
    col = HostArray(abi.I32, vals, N, validity, 0, 0, nc)
    pv = oracle.generate_bits(46, 0, 0.5, N)
    pn = oracle.generate_bits(146, 0, 1 - pred_null_p, N) if pred_null_p is not None else None
    return col, HostArray(BOOL, pv, N, pn, 0, 0, -1 if pn is not None else 0)


def naive_filter(col, pred):
    keep = pred.value_array() & pred.valid_mask()  # a null predicate slot selects nothing (filter.rs:167-171)
    return col.value_array()[keep], col.valid_mask()[keep]


def check(oracle, null_p, pred_null_p):
    col, pred = config1(oracle, null_p, pred_null_p)
    out = oracle.filter(col, pred)
    ev, em = naive_filter(col, pred)
    assert out.length == len(ev) and abs(out.length - N * 0.5 * (1 - (pred_null_p or 0))) < 5000
    assert np.array_equal(out.value_array(), ev)  # bytes under nulls are copied blindly (filter_native ignores validity)
    assert np.array_equal(out.valid_mask(), em)
    if null_p is None or em.all():
        assert out.validity is None  # filter.rs:518-526
    else:
        assert out.null_count == int((~em).sum())
    count, strategy = oracle.filter_plan(pred)
    assert count == len(ev) and strategy == abi.FILTER_INDEX if hasattr(abi, "FILTER_INDEX") else True
    return col, pred, out


def test_config1_filter_int32_1e6_half_selectivity(oracle):
    check(oracle, None, None)


def test_config1_with_nulls(oracle):
    check(oracle, 0.05, None)
    check(oracle, 0.05, 0.05)
    check(oracle, 0.0, None)  # NullBuffer present but no nulls -> result has none
