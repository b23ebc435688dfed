"""Pins the CPU oracle against the reference's own literal test vectors (tests/golden/vectors.json,
transcribed by tests/golden/make_golden.py from the cited arrow-rs test functions)."""
import pytest

from golden_util import load_cases, run_case

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_vector(oracle, case):
    run_case(oracle, case)
