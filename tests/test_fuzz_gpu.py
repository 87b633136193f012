"""A short, fixed-seed slice of the randomised differential test (tests/fuzz_gpu.py): every op on
random shapes / distributions / radii against the oracle, bit-equal."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_cases_match_oracle(dev, seed):
    import fuzz_gpu
    counts, fails = fuzz_gpu.run(seed, 180)
    assert sum(counts.values()) == 180 and len(counts) == 9
    assert not fails, fails
