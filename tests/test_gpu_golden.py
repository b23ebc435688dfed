"""The CUDA path (through the C ABI) against the reference's own literal test vectors."""
import pytest

from golden_util import load_cases, run_case

CASES = load_cases()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_gpu_matches_reference_vector(gpu, case):
    run_case(gpu, case)
