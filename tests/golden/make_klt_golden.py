"""Generates tests/golden/klt_config2.npz — run in the build container where /root/reference and
cv2 4.13.0 exist.  Config 2 of BASELINE.json as constructed in SURVEY.md D8: the reference's
three test images (images/640_480_test.png and its moved / shear variants) converted to gray the
way test/klt_test.cpp:24-26 does, the first 200 FAST(50, nms) corners, and the outputs of
cv2.calcOpticalFlowPyrLK with the reference's exact arguments (KLTTracker.cpp:61-64).
The arrays are inputs/outputs of OpenCV (the real reference arithmetic), not reference source.
"""
import os
import zlib

import cv2
import numpy as np

REF = "/root/reference/images"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "klt_config2.npz")


def gray(name):
    return cv2.cvtColor(cv2.imread(os.path.join(REF, name), cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)


def lk(prev, nxt, pts, init):
    p0 = pts.reshape(-1, 1, 2).copy(); p1 = init.reshape(-1, 1, 2).copy()
    nx, st, er = cv2.calcOpticalFlowPyrLK(prev, nxt, p0, p1, winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW, minEigThreshold=1e-4)
    return nx[:, 0], st[:, 0], er[:, 0]


def main():
    cv2.setNumThreads(1)
    g0, g1, g2 = gray("640_480_test.png"), gray("640_480_moved_test.png"), gray("640_480_shear_test.png")
    fast = cv2.FastFeatureDetector_create(threshold=50, nonmaxSuppression=True)
    kps = fast.detect(g0, None)
    pts_all = np.array([k.pt for k in kps], np.float32)
    rng = np.random.default_rng(0)
    rnd = np.stack([rng.uniform(-5, 645, 700), rng.uniform(-5, 485, 700)], 1).astype(np.float32)
    stress = np.concatenate([pts_all, rnd])
    out = dict(gray0=g0, gray_moved=g1, gray_shear=g2, pts200=pts_all[:200], pts_stress=stress, cv2_version=np.array(cv2.__version__))
    for name, nxt in (("moved", g1), ("shear", g2)):
        for tag, pts in (("200", pts_all[:200]), ("stress", stress)):
            nx, st, er = lk(g0, nxt, pts, pts)
            out[f"{name}_{tag}_next"] = nx; out[f"{name}_{tag}_status"] = st; out[f"{name}_{tag}_err"] = er
    # pyramid levels and Scharr derivatives of the base image straight from OpenCV
    lv = g0
    for l in range(4):
        out[f"pyr{l}"] = lv
        sch = np.stack([cv2.Scharr(lv, cv2.CV_16S, 1, 0), cv2.Scharr(lv, cv2.CV_16S, 0, 1)], -1)
        out[f"scharr{l}_crc32"] = np.array(zlib.crc32(np.ascontiguousarray(sch).tobytes()), np.uint32)
        if l >= 2:  # full arrays only for the small levels (keeps the fixture small)
            out[f"scharr{l}"] = sch
        lv = cv2.pyrDown(lv)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; tracked:", {k: int(v.sum()) for k, v in out.items() if k.endswith("_status")})


if __name__ == "__main__":
    main()
