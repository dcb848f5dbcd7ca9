"""GPU parity tests of the replenishFeatures path (EKFVIO.cpp:224-311) through the C ABI.

Bar: integer work, bit-exact — keypoint set, order and response of cv::FAST; the features accepted by
the greedy scan, in order.  Checked against cv2 4.13.0 golden vectors (tests/golden/make_replenish_golden.py)
and against the oracle (oracle/replenish_oracle.py) on seeded synthetic frames, ragged sizes included.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import replenish_oracle as R  # noqa: E402
import frame_oracle as FO  # noqa: E402

NAMES = ("gray0", "gray_moved", "gray_shear")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "replenish_golden.npz")), np.load(os.path.join(ROOT, "tests", "golden", "klt_config2.npz"))


def test_fast_detect_bit_exact_vs_cv2_golden(cuda, gold):
    import torch
    from ekf_vio_b200 import capi
    G, I = gold
    imgs = np.stack([I[n] for n in NAMES])
    det = capi.FastDetector(640, 480, 3, 8192)
    d = torch.from_numpy(imgs).cuda()
    kp = torch.zeros(3, 8192, 2, dtype=torch.int16, device="cuda"); rs = torch.zeros(3, 8192, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
    det.detect(d, 50, True, kp, rs, cnt); torch.cuda.synchronize()
    for b, n in enumerate(NAMES):
        c = int(cnt[b])
        assert c == len(G[f"{n}_kp"])
        np.testing.assert_array_equal(kp[b, :c].cpu().numpy(), G[f"{n}_kp"])
        np.testing.assert_array_equal(rs[b, :c].cpu().numpy(), G[f"{n}_resp"])
    det.detect(d, 20, False, kp, None, cnt); torch.cuda.synchronize()
    for b, n in enumerate(NAMES):
        c = int(cnt[b])
        np.testing.assert_array_equal(kp[b, :c].cpu().numpy(), G[f"{n}_kp_thr20_nonms"])
    assert det.launches == 6
    det.close()


def test_replenish_h_vs_cv2_golden(cuda, gold):
    from ekf_vio_b200 import capi
    G, I = gold
    imgs = np.stack([I[n] for n in NAMES])
    det = capi.FastDetector(640, 480, 3, 2048)
    K = np.array([[300.0, 0, 320.0], [0, 310.0, 240.0], [0, 0, 1]], np.float32)
    K9 = np.tile(K.T.reshape(-1), (3, 1))
    # empty state, 100 needed
    new_px, metric, n_new, kp, cnt = det.replenish_h(imgs, 50, None, None, np.full(3, 100), K9=K9)
    for b, n in enumerate(NAMES):
        ref = G[f"{n}_new_empty"]
        assert n_new[b] == len(ref)
        np.testing.assert_array_equal(new_px[b, :n_new[b]], ref)
        np.testing.assert_array_equal(metric[b, :n_new[b], 0], ref[:, 0].astype(np.float32) / np.float32(300.0))   # E1: principal point dropped
        np.testing.assert_array_equal(metric[b, :n_new[b], 1], ref[:, 1].astype(np.float32) / np.float32(310.0))
        np.testing.assert_array_equal(kp[b, :cnt[b]], G[f"{n}_kp"])
    # 60 features already in the state (half-integer positions, circles clipped by the border)
    ex = np.stack([G[f"{n}_existing"] for n in NAMES])
    new_px, _, n_new, _, _ = det.replenish_h(imgs, 50, ex, np.full(3, 60), np.full(3, 40))
    for b, n in enumerate(NAMES):
        ref = G[f"{n}_new_60"]
        assert n_new[b] == len(ref)
        np.testing.assert_array_equal(new_px[b, :n_new[b]], ref)
    det.close()


@pytest.mark.parametrize("w,h", [(640, 480), (333, 97), (64, 48), (7, 7), (1280, 720)])
def test_fast_and_scan_vs_oracle_on_synthetic_frames(cuda, w, h):
    """Seeded noise + blocks (many corners), sizes that are not multiples of the tiles, heterogeneous batch."""
    from ekf_vio_b200 import capi
    rng = np.random.default_rng(w * 1000 + h)
    B = 3
    imgs = np.zeros((B, h, w), np.uint8)
    for b in range(B):
        base = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1)).astype(np.uint8)
        img = np.kron(base, np.ones((8, 8), np.uint8))[:h, :w].astype(np.int32)
        img += rng.integers(-12, 13, (h, w))
        imgs[b] = np.clip(img, 0, 255).astype(np.uint8)
    imgs[B - 1] = 128                                                    # a frame without any corner
    cap = 1 << 14
    det = capi.FastDetector(w, h, B, cap)
    nex = np.array([0, 7, 3], np.int32)
    ex = np.stack([rng.uniform(-5, w + 5, (B, 8)), rng.uniform(-5, h + 5, (B, 8))], 2).astype(np.float32)
    needed = np.array([50, 20, 5], np.int32)
    for thr in (30, 5):
        new_px, _, n_new, kp, cnt = det.replenish_h(imgs, thr, ex, nex, needed, min_dist=9, kill_pad=4, max_new=128)
        for b in range(B):
            okp, _ = R.fast9_16(imgs[b], thr, True)
            assert cnt[b] == len(okp)
            stored = min(len(okp), cap)
            np.testing.assert_array_equal(kp[b, :stored], okp[:stored])
            ref = R.select_new_features(okp[:stored], ex[b, :nex[b]], w, h, int(needed[b]), min_dist=9, kill_pad=4)
            assert n_new[b] == min(len(ref), 128)
            np.testing.assert_array_equal(new_px[b, :n_new[b]], ref[:128])
    assert int(cnt[B - 1]) == 0 and int(n_new[B - 1]) == 0
    det.close()


def test_keypoint_capacity_overflow_is_counted_not_stored(cuda, gold):
    from ekf_vio_b200 import capi
    G, I = gold
    det = capi.FastDetector(640, 480, 1, 100)
    new_px, _, n_new, kp, cnt = det.replenish_h(I["gray0"][None], 50, None, None, np.array([100]))
    assert int(cnt[0]) == 666
    np.testing.assert_array_equal(kp[0], G["gray0_kp"][:100])
    ref = R.select_new_features(G["gray0_kp"][:100], [], 640, 480, 100)
    np.testing.assert_array_equal(new_px[0, :n_new[0]], ref)
    det.close()


def test_bad_arguments_fail_loudly(cuda):
    from ekf_vio_b200 import capi
    with pytest.raises(RuntimeError):
        capi.FastDetector(4, 4, 1, 16)
    det = capi.FastDetector(64, 48, 1, 16)
    with pytest.raises(RuntimeError):
        det.replenish_h(np.zeros((2, 48, 64), np.uint8), 50, None, None, np.array([1, 1]))   # batch above capacity
    det.close()


def test_frame_resize_bit_exact(cuda, gold):
    """Frame::Frame's cv::resize (Frame.cpp:19): cv2 golden at 2x (area-fast), 3x, 4x (the default), 5x; oracle on ragged sizes."""
    import torch
    from ekf_vio_b200 import capi
    G, I = gold
    g0 = I["gray0"]
    d = torch.from_numpy(np.stack([g0, I["gray_moved"]])).cuda()
    for s in (1, 2, 3, 4, 5):
        out = capi.frame_resize(d, s).cpu().numpy()
        ref = g0 if s == 1 else G[f"gray0_resize{s}"]
        np.testing.assert_array_equal(out[0], ref)
        np.testing.assert_array_equal(out[1], FO.resize(I["gray_moved"], s))
    crop = torch.from_numpy(np.ascontiguousarray(g0[:479, :639])[None]).cuda()
    np.testing.assert_array_equal(capi.frame_resize(crop, 2).cpu().numpy()[0], G["gray0_crop_resize2"])
    rng = np.random.default_rng(5)
    for (w, h) in [(752, 480), (641, 479), (97, 33), (1280, 720)]:
        img = rng.integers(0, 256, (3, h, w)).astype(np.uint8)
        dd = torch.from_numpy(img).cuda()
        for s in (2, 3, 4, 7):
            out = capi.frame_resize(dd, s).cpu().numpy()
            for b in range(3):
                np.testing.assert_array_equal(out[b], FO.resize(img[b], s))
