"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of Frame::Frame (Frame.cpp:15-41): cv::resize(img, scaled, Size(cols/inv_scale, rows/inv_scale))
with the default INTER_LINEAR on 8-bit images, and the K scaling.  OpenCV (third-party, unpinned in the
reference) is restated from its published algorithm (imgproc/resize.cpp: area-fast substitution at exactly
2x, 11-bit fixed-point HResizeLinear / VResizeLinear otherwise) and pinned against the in-container cv2 4.13.0
by tests/test_oracle_replenish.py and tests/golden/replenish_golden.npz.
"""
import numpy as np


def _coeffs(dn: int, sn: int):
    scale = 1.0 / (dn / sn)                       # resize.cpp: scale_x = 1. / inv_scale_x
    idx = np.zeros(dn, np.int64); a = np.zeros((dn, 2), np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f)); f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= sn - 1:
            s, f = sn - 1, np.float32(0)
        idx[d] = s
        a[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        a[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a


def resize(img: np.ndarray, inv_scale: int) -> np.ndarray:
    sh, sw = img.shape
    dw, dh = sw // inv_scale, sh // inv_scale
    if inv_scale == 1:
        return img.copy()
    S = img.astype(np.int64)
    if inv_scale == 2 and dw * 2 == sw and dh * 2 == sh:
        return ((S[0::2, 0::2] + S[0::2, 1::2] + S[1::2, 0::2] + S[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xi, xa = _coeffs(dw, sw); yi, ya = _coeffs(dh, sh)
    x1 = np.minimum(xi + 1, sw - 1); y1 = np.minimum(yi + 1, sh - 1)
    rows = S[:, xi] * xa[:, 0] + S[:, x1] * xa[:, 1]
    out = (((ya[:, 0:1] * (rows[yi] >> 4)) >> 16) + ((ya[:, 1:2] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def scale_K(k9_row_major, inv_scale: int) -> np.ndarray:
    """Frame.cpp:24-30: K(0,0), K(0,2), K(1,1), K(1,2) divided by inv_scale, K(2,2) = 1, rest 0 (float)."""
    k = np.asarray(k9_row_major, np.float64)
    K = np.zeros((3, 3), np.float32)
    K[0, 0] = k[0] / inv_scale; K[0, 2] = k[2] / inv_scale; K[1, 1] = k[4] / inv_scale; K[1, 2] = k[5] / inv_scale; K[2, 2] = 1.0
    return K
