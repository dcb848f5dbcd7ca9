"""CPU tests of the KLT oracle against golden vectors produced by OpenCV 4.13.0 (and against cv2
itself when importable).  No GPU."""
import os
import zlib

import numpy as np
import pytest

from tests import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "klt_config2.npz"))


def test_pyramid_and_scharr_bit_exact_vs_cv2_golden():
    g0 = GOLD["gray0"]
    for l in range(4):
        img, der, ml = O.klt_level(g0, 21, 3, l)
        assert ml == 3
        np.testing.assert_array_equal(img, GOLD[f"pyr{l}"])
        assert zlib.crc32(np.ascontiguousarray(der).tobytes()) == int(GOLD[f"scharr{l}_crc32"])
        if l >= 2:
            np.testing.assert_array_equal(der, GOLD[f"scharr{l}"])


def test_pyramid_stops_when_next_level_not_larger_than_window():
    lw = (O.C.c_int * 16)(); lh = (O.C.c_int * 16)()
    O.klt.klt_oracle_level_sizes.restype = O.C.c_int
    assert O.klt.klt_oracle_level_sizes(640, 480, 21, 3, lw, lh) == 3 and (lw[3], lh[3]) == (80, 60)
    assert O.klt.klt_oracle_level_sizes(64, 32, 21, 3, lw, lh) == 0
    assert O.klt.klt_oracle_level_sizes(160, 120, 21, 3, lw, lh) == 2 and (lw[2], lh[2]) == (40, 30)
    assert O.klt.klt_oracle_level_sizes(87, 45, 21, 3, lw, lh) == 1     # 44x23 > 21, 22x12 not


@pytest.mark.parametrize("pair,tracked", [("moved", 190), ("shear", 194)])
def test_config2_tracking_vs_cv2_golden(pair, tracked):
    g0, g1, pts = GOLD["gray0"], GOLD[f"gray_{pair}"], GOLD["pts200"]
    on, ost, oer, it = O.klt_calc_optical_flow(g0, g1, pts, pts)
    np.testing.assert_array_equal(ost, GOLD[f"{pair}_200_status"])
    assert int(ost.sum()) == tracked
    ok = ost == 1
    assert np.abs(on[ok] - GOLD[f"{pair}_200_next"][ok]).max() <= 1e-3
    assert np.abs(oer[ok] - GOLD[f"{pair}_200_err"][ok]).max() <= 5e-3
    assert 10 < it.mean() < 18                              # about 13-14 LK iterations per feature over 4 levels


@pytest.mark.parametrize("pair", ["moved", "shear"])
def test_stress_points_status_identical(pair):
    g0, g1, pts = GOLD["gray0"], GOLD[f"gray_{pair}"], GOLD["pts_stress"]
    on, ost, _, _ = O.klt_calc_optical_flow(g0, g1, pts, pts)
    np.testing.assert_array_equal(ost, GOLD[f"{pair}_stress_status"])
    ref = GOLD[f"{pair}_stress_next"]
    pad = ~((ref[:, 0] < 11) | (ref[:, 1] < 11) | (640 - ref[:, 0] < 11) | (480 - ref[:, 1] < 11))
    ok = (ost == 1) & pad
    assert np.abs(on[:666][ok[:666]] - ref[:666][ok[:666]]).max() <= 0.01


def test_odd_sizes_vs_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for (w, h) in [(77, 101), (333, 245), (641, 479)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        d = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
        O.klt.klt_oracle_pyrdown_u8(O.P(img), w, h, w, O.P(d), d.shape[1])
        np.testing.assert_array_equal(d, cv2.pyrDown(img))
        s = np.zeros((h, w, 2), np.int16)
        O.klt.klt_oracle_scharr_s16(O.P(img), w, h, w, O.P(s))
        np.testing.assert_array_equal(s[..., 0], cv2.Scharr(img, cv2.CV_16S, 1, 0))
        np.testing.assert_array_equal(s[..., 1], cv2.Scharr(img, cv2.CV_16S, 0, 1))


def test_postprocess_epilogue_and_k_index_error():
    """KLTTracker.cpp:72-92 with Feature::pixel2Metric's linear K indexing (E1)."""
    K = np.zeros((3, 3), np.float32); K[0, 0] = 400; K[1, 1] = 410; K[0, 2] = 320; K[1, 2] = 240; K[2, 2] = 1
    K9 = np.ascontiguousarray(K.T).reshape(9)
    pts = np.array([[100.5, 200.25], [5.0, 100.0], [630.0, 100.0], [100.0, 470.0], [300.0, 300.0]], np.float32)
    st = np.array([1, 1, 1, 1, 0], np.uint8)
    m, c, p = O.klt_postprocess(pts, st, 640, 480, K9)
    assert list(p) == [1, 0, 0, 0, 0]
    np.testing.assert_allclose(m[0], [100.5 / 400, 200.25 / 410], rtol=1e-6)      # principal point NOT subtracted
    np.testing.assert_allclose(c[0], [1e-5 / 400 ** 2, 0, 0, 1e-5 / 410 ** 2], rtol=1e-5)
    assert not c[1:].any()
