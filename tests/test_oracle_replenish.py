"""CPU tests: the replenishFeatures oracle (oracle/replenish_oracle.py) against golden vectors produced by
cv2 4.13.0 (tests/golden/make_replenish_golden.py) and against cv2 itself when importable.  No GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import replenish_oracle as R  # noqa: E402
import frame_oracle as FO  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "replenish_golden.npz"))
IMGS = np.load(os.path.join(ROOT, "tests", "golden", "klt_config2.npz"))
NAMES = ("gray0", "gray_moved", "gray_shear")


@pytest.mark.parametrize("name", NAMES)
def test_fast_matches_cv2_golden(name):
    kp, resp = R.fast9_16(IMGS[name], 50, True)
    np.testing.assert_array_equal(kp, GOLD[f"{name}_kp"])          # set and order
    np.testing.assert_array_equal(resp, GOLD[f"{name}_resp"])      # cornerScore
    kp2, _ = R.fast9_16(IMGS[name], 20, False)
    np.testing.assert_array_equal(kp2, GOLD[f"{name}_kp_thr20_nonms"])


@pytest.mark.parametrize("name", NAMES)
def test_greedy_scan_matches_cv2_golden(name):
    kp = GOLD[f"{name}_kp"]
    np.testing.assert_array_equal(R.select_new_features(kp, [], 640, 480, 100), GOLD[f"{name}_new_empty"])
    np.testing.assert_array_equal(R.select_new_features(kp, GOLD[f"{name}_existing"], 640, 480, 40), GOLD[f"{name}_new_60"])


def test_filled_circle_matches_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(200):
        cx, cy, r = int(rng.integers(-40, 680)), int(rng.integers(-40, 520)), int(rng.integers(0, 45))
        a = np.zeros((480, 640), np.uint8); cv2.circle(a, (cx, cy), r, 255, -1)
        b = np.zeros((480, 640), np.uint8); R.draw_filled_circle(b, cx, cy, r)
        np.testing.assert_array_equal(a, b)


def test_pixel2metric_uses_linear_indices_of_K():
    """E1: K(2) and K(5) of a column-major 3x3 are K(2,0) and K(2,1) (= 0), not the principal point."""
    K = np.array([[300.0, 0, 320.0], [0, 310.0, 240.0], [0, 0, 1]], np.float32)
    kp = np.array([[100, 50]], np.int32)
    px, metric = R.select_new_features(kp, [], 640, 480, 1, K9=K.T.reshape(-1))
    np.testing.assert_allclose(metric[0], [100 / 300.0, 50 / 310.0], rtol=1e-6)


def test_frame_resize_matches_cv2_golden():
    g0 = IMGS["gray0"]
    for s in (2, 3, 4, 5):
        np.testing.assert_array_equal(FO.resize(g0, s), GOLD[f"gray0_resize{s}"])
    np.testing.assert_array_equal(FO.resize(np.ascontiguousarray(g0[:479, :639]), 2), GOLD["gray0_crop_resize2"])
    np.testing.assert_array_equal(FO.resize(g0, 1), g0)


def test_frame_resize_matches_cv2_on_random_sizes_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (w, h) in [(752, 480), (1280, 720), (641, 479), (97, 33)]:
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        for s in (2, 3, 4, 7):
            np.testing.assert_array_equal(FO.resize(img, s), cv2.resize(img, (w // s, h // s)))
