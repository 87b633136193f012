"""The algorithmic-byte formulas behind `roofline.achieved` (pointnet2_b200/workloads.py) against the
worked examples of SURVEY.md §8(d), and the synthetic input recipes' basic properties."""
import numpy as np
import pytest

from pointnet2_b200 import workloads as W


def test_algorithmic_bytes_match_the_survey_worked_examples():
    b, n, m, s = 32, 4096, 1024, 32  # cfg2
    assert W.bytes_fps(b, n, m, with_new_xyz=False) == 12 * b * n + 4 * b * m == 1_703_936          # 1.70 MB
    assert W.bytes_fps(b, n, m, with_new_xyz=True) == 12 * b * n + 16 * b * m == 2_097_152            # + fused new_xyz
    assert W.bytes_ball_query(b, n, m, s) == 12 * b * n + 12 * b * m + 4 * b * m * s + 4 * b * m == 6_291_456
    assert W.bytes_group(b, n, m, s, 3) == 4 * b * m * s + 4 * b * min(n, m * s) * 3 + 4 * b * m * s * 3 == 18_350_080
    assert W.bytes_sa_layer(b, n, m, s) == 27_262_976                                                  # 27.3 MB per layer
    # cfg3 layer 2, C=320: 189.3 / 357.6 / 694.2 MB for S = 32 / 64 / 128
    assert [W.bytes_group(32, 512, 128, k, 320) for k in (32, 64, 128)] == [189_267_968, 357_564_416, 694_157_312]


@pytest.mark.parametrize("gen", ["U", "S", "D"])
def test_input_recipes_are_seeded_finite_float32(gen):
    a = W.DISTRIBUTIONS[gen](3, 500, 7)
    b2 = W.DISTRIBUTIONS[gen](3, 500, 7)
    assert a.dtype == np.float32 and a.shape == (3, 500, 3) and np.isfinite(a).all()
    np.testing.assert_array_equal(a, b2)
    assert not np.array_equal(a, W.DISTRIBUTIONS[gen](3, 500, 8))
    if gen == "S":
        assert np.abs(a).max() <= 1.0 + 1e-6            # pc_normalize: unit sphere
    if gen == "D":
        assert len(np.unique(a[0], axis=0)) < 0.5 * 500  # duplicate-heavy: FPS ties are real


def test_baseline_configs_are_the_survey_configs():
    c = W.CFG2_SSG_SA
    assert (c["b"], c["n"], c["npoint"], c["nsample"], c["radius"], c["dist"]) == (32, 4096, 1024, 32, 0.1, "U")
    c1 = W.CFG1_FPS_CPU
    assert (c1["b"], c1["n"], c1["npoint"]) == (8, 1024, 512)
