"""Generates tests/golden/replenish_golden.npz with the in-container cv2 (4.13.0): the reference's
replenishFeatures arithmetic (EKFVIO.cpp:224-311) is cv::FAST + cv::circle + a scalar scan, so the
golden vectors are produced by exactly those OpenCV calls.  Run from the repo root:
    python tests/golden/make_replenish_golden.py
Inputs are the committed 640x480 test images of klt_config2.npz (the reference's images/640_480_test*.png)."""
import numpy as np, cv2

g = np.load("tests/golden/klt_config2.npz")
out = {}
W, H, THR, MIN_DIST, KILL_PAD, NUM_FEATURES = 640, 480, 50, 30, 11, 100   # Params.h:24,43,33,46


def reference_scan(img, existing_px, needed):
    det = cv2.FastFeatureDetector_create(threshold=THR, nonmaxSuppression=True)
    kp = det.detect(img)
    check = np.zeros(img.shape, np.uint8)
    for ex, ey in existing_px:
        cv2.circle(check, (int(np.rint(np.float32(ex))), int(np.rint(np.float32(ey)))), MIN_DIST, 255, -1)
    new = []
    i = 0
    while i < needed and i < len(kp):
        x, y = int(kp[i].pt[0]), int(kp[i].pt[1])
        if check[y, x]:
            needed += 1
        elif x < KILL_PAD or y < KILL_PAD or W - x < KILL_PAD or H - y < KILL_PAD:
            needed += 1
        else:
            cv2.circle(check, (x, y), MIN_DIST, 255, -1)
            new.append((x, y))
        i += 1
    kps = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kp], np.int16)
    resp = np.array([int(k.response) for k in kp], np.int32)
    return kps, resp, np.array(new, np.int16).reshape(-1, 2)


rng = np.random.default_rng(7)
for name in ("gray0", "gray_moved", "gray_shear"):
    img = g[name]
    # (a) start-up: empty state, START_FEATURE_COUNT..NUM_FEATURES needed; (b) 60 features already tracked
    kps, resp, new_a = reference_scan(img, [], NUM_FEATURES)
    existing = np.stack([rng.uniform(0, W, 60), rng.uniform(0, H, 60)], 1).astype(np.float32)
    existing[:5] += np.float32(0.5)          # exact .5 fractions exercise cvRound's half-to-even
    existing[5] = (-12.0, 100.0); existing[6] = (650.0, 470.0)   # circles clipped by the image border
    _, _, new_b = reference_scan(img, existing, NUM_FEATURES - 60)
    out[f"{name}_kp"] = kps; out[f"{name}_resp"] = resp
    out[f"{name}_new_empty"] = new_a
    out[f"{name}_existing"] = existing; out[f"{name}_new_60"] = new_b
    for thr in (20,):
        det = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=False)
        kp2 = det.detect(img)
        out[f"{name}_kp_thr{thr}_nonms"] = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kp2], np.int16)
    print(name, len(kps), "keypoints;", len(new_a), "new from empty;", len(new_b), "new with 60 existing")
# Frame::Frame's cv::resize (Frame.cpp:19) at the reference's default INVERSE_IMAGE_SCALE 4 and at 2, 3, 5
for s in (2, 3, 4, 5):
    out[f"gray0_resize{s}"] = cv2.resize(g["gray0"], (W // s, H // s))
out["gray0_crop_resize2"] = cv2.resize(np.ascontiguousarray(g["gray0"][:479, :639]), (639 // 2, 479 // 2))   # 2x but not area-fast
np.savez_compressed("tests/golden/replenish_golden.npz", **out)
