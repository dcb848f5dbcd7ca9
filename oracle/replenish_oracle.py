"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of EKFVIO::replenishFeatures (EKFVIO.cpp:224-311):
  cv::FAST(img, kp, FAST_THRESHOLD, nonmaxSuppression=true)    (:242)   -> fast9_16()
  cv::circle(checkImg, centre, MIN_NEW_FEATURE_DIST, 255, -1)   (:258-260, :297) -> filled_circle_spans()
  the greedy scan over the keypoints in detector order          (:262-305) -> select_new_features()
  Frame::isPixelInBox (Frame.cpp:44-55), Feature::pixel2Metric (Feature.h:60-62, E1: K(2), K(5) linear indices)

OpenCV is a third-party dependency of the reference (unpinned); the restatement follows the
published algorithm of cv::FAST (FAST-9/16, Rosten & Drummond; OpenCV's cornerScore<16> for the response)
and of cv::circle's midpoint rasteriser, and is pinned against the in-container cv2 4.13.0 by
tests/test_oracle_replenish.py and the golden vectors of tests/golden/make_replenish_golden.py.
"""
import numpy as np

# the 16-pixel Bresenham ring of radius 3, OpenCV order (x, y)
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
        (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_scores(img: np.ndarray, threshold: int) -> np.ndarray:
    """Score image: 0 where the pixel is not a FAST-9/16 corner, else cornerScore<16>() + 1.
    A corner has >= 9 contiguous ring pixels all > v + t or all < v - t (strict); the 3-pixel border is skipped.
    cornerScore = max(t, max over the 16 arcs of min(d), max over arcs of min(-d)) - 1 with d = v - ring."""
    h, w = img.shape
    v = img.astype(np.int32)
    d = np.zeros((16, h, w), np.int32)
    for k, (dx, dy) in enumerate(RING):
        sh = np.zeros_like(v)
        ys, ye = max(0, -dy), min(h, h - dy)
        xs, xe = max(0, -dx), min(w, w - dx)
        sh[ys:ye, xs:xe] = v[ys + dy:ye + dy, xs + dx:xe + dx]
        d[k] = v - sh
    darker = d > threshold            # ring pixel darker than the centre by more than t
    brighter = d < -threshold
    is_corner = np.zeros((h, w), bool)
    best_pos = np.full((h, w), -(1 << 30), np.int32)   # max over arcs of min(d)
    best_neg = np.full((h, w), -(1 << 30), np.int32)   # max over arcs of min(-d)
    for s in range(16):
        idx = [(s + j) % 16 for j in range(9)]
        is_corner |= np.all(darker[idx], axis=0) | np.all(brighter[idx], axis=0)
        best_pos = np.maximum(best_pos, d[idx].min(axis=0))
        best_neg = np.maximum(best_neg, (-d[idx]).min(axis=0))
    score = np.maximum(np.maximum(best_pos, best_neg), threshold) - 1
    out = np.where(is_corner, score + 1, 0).astype(np.int32)
    out[:3, :] = 0; out[-3:, :] = 0; out[:, :3] = 0; out[:, -3:] = 0
    return out


def fast9_16(img: np.ndarray, threshold: int, nonmax: bool = True):
    """Keypoints (x, y) in OpenCV's order (row by row, left to right) and their responses."""
    s = fast_scores(img, threshold)
    keep = s > 0
    if nonmax:
        p = np.pad(s, 1)
        h, w = s.shape
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == 0 and dy == 0:
                    continue
                keep &= s > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    ys, xs = np.nonzero(keep)        # row-major order
    return np.stack([xs, ys], 1).astype(np.int32), (s[ys, xs] - 1).astype(np.int32)


def filled_circle_spans(cx: int, cy: int, radius: int):
    """Row spans (y, x_first, x_last) of cv::circle(..., thickness=-1): OpenCV's midpoint rasteriser
    (drawing.cpp, Circle()).  Spans may repeat rows and may lie outside the image (the caller clips)."""
    spans = []
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        spans.append((cy - dy, cx - dx, cx + dx)); spans.append((cy + dy, cx - dx, cx + dx))
        spans.append((cy - dx, cx - dy, cx + dy)); spans.append((cy + dx, cx - dy, cx + dy))
        dy += 1
        err += plus; plus += 2
        mask = -1 if err > 0 else 0           # (err <= 0) - 1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return spans


def draw_filled_circle(mask: np.ndarray, cx: int, cy: int, radius: int):
    h, w = mask.shape
    for y, x0, x1 in filled_circle_spans(cx, cy, radius):
        if 0 <= y < h:
            a, b = max(x0, 0), min(x1, w - 1)
            if a <= b:
                mask[y, a:b + 1] = 255


def cv_round(x: float) -> int:
    """cvRound: round half to even (Point2f -> Point conversion)."""
    return int(np.rint(np.float64(np.float32(x))))


def select_new_features(kps, existing_px, width, height, needed, min_dist=30, kill_pad=11, K9=None):
    """The greedy scan of EKFVIO.cpp:262-305.  kps: (x, y) in detector order; existing_px: float pixel
    positions of the features already in the state.  Returns the accepted pixels (in order) and, when a
    column-major 3x3 K9 is given, their metric coordinates (pixel2Metric with the linear-index E1 semantics)."""
    mask = np.zeros((height, width), np.uint8)
    for ex, ey in existing_px:
        draw_filled_circle(mask, cv_round(ex), cv_round(ey), int(min_dist))
    out = []
    i = 0
    while i < needed and i < len(kps):
        x, y = int(kps[i][0]), int(kps[i][1])
        if mask[y, x]:
            needed += 1
        elif x < kill_pad or y < kill_pad or width - x < kill_pad or height - y < kill_pad:
            needed += 1
        else:
            draw_filled_circle(mask, x, y, int(min_dist))
            out.append((x, y))
        i += 1
    px = np.array(out, np.int32).reshape(-1, 2)
    if K9 is None:
        return px
    K9 = np.asarray(K9, np.float32)
    metric = np.stack([(px[:, 0].astype(np.float32) - K9[2]) / K9[0], (px[:, 1].astype(np.float32) - K9[5]) / K9[4]], 1)
    return px, metric
