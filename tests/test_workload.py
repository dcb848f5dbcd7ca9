"""Host logic: the synthetic workload is a pure function of the GLOBAL filter / sequence index, so
sharding across ranks changes nothing; flop/byte accounting of bench.py matches SURVEY.md §8d."""
import numpy as np

import bench
from ekf_vio_b200 import workload


def test_ekf_streams_are_sharding_invariant():
    uv, meas, truth = workload.ekf_streams(0, 8, 5, 4)
    uv_b, meas_b, truth_b = workload.ekf_streams(4, 4, 5, 4)
    np.testing.assert_array_equal(uv[4:], uv_b)
    np.testing.assert_array_equal(meas[:, 4:], meas_b)
    np.testing.assert_array_equal(truth[:, 4:], truth_b)
    assert np.isfinite(meas).all() and np.abs(meas).max() < 3.0


def test_ekf_streams_are_consistent_projections():
    uv, meas, truth = workload.ekf_streams(3, 2, 6, 10, dt=0.05)
    # first measurement is close to the initial projection (small motion), later ones drift smoothly
    assert np.abs(meas[0] - uv).max() < 0.05
    assert np.abs(np.diff(meas, axis=0)).max() < 0.05
    np.testing.assert_allclose(np.linalg.norm(truth[:, :, 3:7], axis=-1), 1.0, atol=1e-12)


def test_klt_pairs_deterministic_and_textured():
    a = workload.klt_pairs(5, 2, 160, 120, 16)
    b = workload.klt_pairs(6, 1, 160, 120, 16)
    np.testing.assert_array_equal(a[0][1], b[0][0]); np.testing.assert_array_equal(a[1][1], b[1][0])
    assert a[0].dtype == np.uint8 and a[0].std() > 20


def test_flop_and_byte_accounting_matches_survey():
    assert abs(bench.flops_per_filter_step(30) / 1e6 - 4.60) < 0.02
    assert abs(bench.flops_per_filter_step(50) / 1e6 - 17.24) < 0.02
    assert abs(bench.flops_per_filter_step(300) / 1e6 - 2822) < 2
    assert bench.KLT_BYTES_WITH_DERIVS == 408000 + 100800 + 1632000 and bench.KLT_BYTES_NO_DERIVS == 408000 + 100800


def test_vio_sequences_are_pure_functions_of_the_sequence_index():
    """Sharding the frame loop over ranks must not change any sequence: frames depend on the global index only."""
    from ekf_vio_b200 import workload
    a = workload.vio_sequences(0, 4, 3, 160, 120)
    b = workload.vio_sequences(2, 2, 3, 160, 120)
    assert a.shape == (3, 4, 120, 160) and a.dtype == np.uint8
    np.testing.assert_array_equal(a[:, 2:], b)
    assert (a[0, 0] != a[1, 0]).any() and a.std() > 5         # the camera pans over a textured scene


def test_frame_oracle_scales_K_like_the_reference():
    """Frame.cpp:24-30: fx, cx, fy, cy divided by inv_scale, K(2,2) = 1."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import frame_oracle as FO
    K = FO.scale_K([400.0, 0, 320.0, 0, 410.0, 240.0, 0, 0, 1.0], 4)
    np.testing.assert_allclose(K, np.array([[100.0, 0, 80.0], [0, 102.5, 60.0], [0, 0, 1.0]], np.float32))
