"""SlicesIterator (arrow-select/src/filter.rs:44-77) on the CPU oracle, pinned on the reference's own cases
(test_slice_iterator_bits / _bits1 / _chunk_and_bits, filter.rs:1640-1678) and against a numpy restatement."""
import numpy as np

from acu import HostArray


def expected_runs(bits):
    runs, start = [], None
    for i, b in enumerate(bits):
        if b and start is None:
            start = i
        if not b and start is not None:
            runs.append((start, i))
            start = None
    if start is not None:
        runs.append((start, len(bits)))
    return runs


def test_reference_cases(oracle):
    f = HostArray.bool_from_numpy(np.array([i == 1 for i in range(64)]))
    assert oracle.filter_slices(f) == [(1, 2)]                                        # filter.rs:1641-1651
    f = HostArray.bool_from_numpy(np.array([i != 1 for i in range(64)]))
    assert oracle.filter_slices(f) == [(0, 1), (2, 64)]                               # :1654-1664
    f = HostArray.bool_from_numpy(np.array([i % 62 != 0 for i in range(130)]))
    assert oracle.filter_slices(f) == [(1, 62), (63, 124), (125, 130)]                # :1667-1677


def test_fuzz_with_offsets_and_nulls(oracle):
    rng = np.random.default_rng(3)
    for n in [0, 1, 63, 64, 65, 129, 1000, 4097]:
        for p in (0.0, 0.05, 0.5, 0.95, 1.0):
            bits = rng.random(n) < p
            mask = rng.random(n) >= 0.1
            f = HostArray.bool_from_numpy(bits, mask, bit_offset=int(rng.integers(0, 9)), mask_offset=int(rng.integers(0, 9)))
            assert oracle.filter_slices(f) == expected_runs(bits & mask)
