"""Synthetic workloads for bench.py and the size-independent tests (numpy; cv2 only to warp images where it is installed).

EKF streams follow the reference's own simulation harness (test/analyzeEKFSimulation.cpp:10-125):
landmarks at depth ~0.5 m in front of the camera, a constant body-frame velocity / angular rate
trajectory, exact projections as measurements, R = 1e-5 I, every feature measured — generated per
filter from a counter-based RNG keyed on the *global* filter index, so a filter gets the same
stream whichever rank it lands on.  KLT pairs are a smooth random
texture and a sub-pixel translated / sheared copy of it.
"""
from __future__ import annotations

import numpy as np


def _qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx], -1)


def _qrot(q, v):
    qv = q[..., 1:]
    uv = 2.0 * np.cross(qv, v)
    return v + q[..., :1] * uv + np.cross(qv, uv)


def _qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def ekf_streams(first_filter: int, num_filters: int, n: int, steps: int, dt: float = 0.05, depth_sigma: float = 0.01,
                vel_range: float = 0.2, omega_range: float = 0.2):
    """Returns init_uv [F,n,2], meas [steps,F,n,2], truth [steps,F,22] (float64).

    Filter g = first_filter + i draws from Philox(key=g): depth z = 0.5 + N(0, depth_sigma),
    x, y = U(-1.5, 1.5) * z, body velocity ~ U(-vel_range, vel_range)^3, omega likewise, a = 0
    (defaults = SURVEY.md §8d config 3: U(-0.2, 0.2) m/s and rad/s, depth sigma 0.01).
    """
    F = num_filters
    X = np.zeros((F, n, 3)); vel = np.zeros((F, 3)); om = np.zeros((F, 3))
    for i in range(F):
        rng = np.random.Generator(np.random.Philox(key=first_filter + i))
        z = 0.5 + rng.normal(0.0, depth_sigma, n)
        X[i, :, 2] = z
        X[i, :, 0] = rng.uniform(-1.5, 1.5, n) * z
        X[i, :, 1] = rng.uniform(-1.5, 1.5, n) * z
        vel[i] = rng.uniform(-vel_range, vel_range, 3)
        om[i] = rng.uniform(-omega_range, omega_range, 3)
    init_uv = X[:, :, :2] / X[:, :, 2:3]
    pos = np.zeros((F, 3)); quat = np.zeros((F, 4)); quat[:, 0] = 1.0
    on = np.linalg.norm(om, axis=1, keepdims=True)
    theta = dt * on
    safe = np.where(on > 0, on, 1.0)
    dq = np.concatenate([np.cos(theta / 2), om / safe * np.sin(theta / 2)], 1)
    dqi = _qconj(dq)
    meas = np.zeros((steps, F, n, 2)); truth = np.zeros((steps, F, 22))
    v = vel.copy()
    for s in range(steps):
        pos = pos + _qrot(quat, dt * v)
        v = _qrot(dqi, v)
        quat = _qmul(quat, dq)
        qi = _qconj(quat)
        rel = X - pos[:, None, :]
        fp = _qrot(qi[:, None, :], rel)
        meas[s] = fp[:, :, :2] / fp[:, :, 2:3]
        truth[s, :, 0:3] = pos; truth[s, :, 3:7] = quat; truth[s, :, 7:10] = v; truth[s, :, 10:13] = om
    return init_uv, meas, truth


def texture(rng: np.random.Generator, h: int, w: int, cell: int = 6) -> np.ndarray:
    """Smooth random texture with corners at every scale the pyramid sees (float64, 0..255)."""
    out = np.zeros((h, w))
    for c, amp in ((cell, 1.0), (cell * 3, 0.7), (cell * 8, 0.5)):
        gh, gw = h // c + 3, w // c + 3
        g = rng.uniform(0, 1, (gh, gw))
        g = (g > 0.5).astype(np.float64) * 0.7 + g * 0.3
        up = np.kron(g, np.ones((c, c)))[:h + 2 * c, :w + 2 * c]
        k = np.ones(c) / c
        up = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, up)
        up = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, up)
        out += amp * up[c:c + h, c:c + w]
    out -= out.min()
    return out / out.max() * 255.0


def _bilinear(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    h, w = img.shape
    xs = np.clip(xs, 0, w - 1.001); ys = np.clip(ys, 0, h - 1.001)
    x0 = np.floor(xs).astype(int); y0 = np.floor(ys).astype(int)
    a = xs - x0; b = ys - y0
    return ((1 - a) * (1 - b) * img[y0, x0] + a * (1 - b) * img[y0, x0 + 1] + (1 - a) * b * img[y0 + 1, x0] + a * b * img[y0 + 1, x0 + 1])


def klt_pairs(first_seq: int, num_pairs: int, w: int = 640, h: int = 480, npts: int = 200, max_shift: float = 8.0, max_shear: float = 0.02):
    """Returns prev [B,h,w] u8, next [B,h,w] u8, pts [B,npts,2] f32, flow [B,2] f32 (mean translation).

    Pair g = first_seq + i: texture from Philox(key=g), next(x, y) = prev(x - tx - s*y, y - ty)."""
    prev = np.zeros((num_pairs, h, w), np.uint8); nxt = np.zeros_like(prev)
    pts = np.zeros((num_pairs, npts, 2), np.float32); flow = np.zeros((num_pairs, 2), np.float32)
    m = 48
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    textures = {}
    for i in range(num_pairs):
        g = first_seq + i
        rng = np.random.Generator(np.random.Philox(key=1_000_003 + g))
        tkey = g % 8                                   # 8 distinct scenes; the motion differs per pair
        if tkey not in textures:
            textures[tkey] = texture(np.random.Generator(np.random.Philox(key=7_000_001 + tkey)), h + 2 * m, w + 2 * m)
        tex = textures[tkey]
        tx, ty = rng.uniform(-max_shift, max_shift, 2)
        sh = rng.uniform(-max_shear, max_shear)
        prev[i] = np.clip(np.rint(tex[m:m + h, m:m + w]), 0, 255).astype(np.uint8)
        xs = xx - tx - sh * (yy - h / 2) + m
        ys = yy - ty + m
        nxt[i] = np.clip(np.rint(_bilinear(tex, xs, ys)), 0, 255).astype(np.uint8)
        pts[i, :, 0] = rng.uniform(40, w - 40, npts)
        pts[i, :, 1] = rng.uniform(40, h - 40, npts)
        flow[i] = (tx, ty)
    return prev, nxt, pts, flow


_GOLDEN = None


def config2_fixture():
    """BASELINE.json configs[1]: gray = cvtColor(imread(images/640_480_test.png)), its moved / shear
    variants and the first 200 FAST(50, nms) corners, from the committed fixture tests/golden/klt_config2.npz (made by
    tests/golden/make_klt_golden.py from the reference's images; /root/reference is not needed at run time)."""
    global _GOLDEN
    if _GOLDEN is None:
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "klt_config2.npz")
        d = np.load(path)
        _GOLDEN = {k: d[k] for k in ("gray0", "gray_moved", "gray_shear", "pts200", "moved_200_status", "shear_200_status")}
    return _GOLDEN


def _warp_affine_reflect101(img: np.ndarray, M: np.ndarray) -> np.ndarray:
    """dst = warpAffine(img, M, INTER_LINEAR, BORDER_REFLECT_101): cv2 where it is installed (the GPU box has it), else numpy."""
    try:
        import cv2
        return cv2.warpAffine(img, M.astype(np.float64), (img.shape[1], img.shape[0]), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    except ImportError:
        h, w = img.shape
        A = np.vstack([M, [0, 0, 1]]); Ai = np.linalg.inv(A)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        xs = Ai[0, 0] * xx + Ai[0, 1] * yy + Ai[0, 2]; ys = Ai[1, 0] * xx + Ai[1, 1] * yy + Ai[1, 2]
        x0 = np.floor(xs).astype(int); y0 = np.floor(ys).astype(int); a = xs - x0; b = ys - y0

        def at(yi, xi):
            xi = np.abs(xi); xi = np.where(xi >= w, 2 * (w - 1) - xi, xi); yi = np.abs(yi); yi = np.where(yi >= h, 2 * (h - 1) - yi, yi)
            return img[np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)].astype(np.float64)
        v = (1 - a) * (1 - b) * at(y0, x0) + a * (1 - b) * at(y0, x0 + 1) + (1 - a) * b * at(y0 + 1, x0) + a * b * at(y0 + 1, x0 + 1)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def klt_pairs_8d(first_seq: int, num_pairs: int, max_shift: float = 24.0, max_shear: float = 0.05):
    """SURVEY.md §8d "additional synthetic sequences for rate": the base image of config 2 translated / sheared by a seeded random
    affine map (|t| <= 24 px, |shear| <= 0.05, seed = sequence index) with warpAffine(INTER_LINEAR, BORDER_REFLECT_101); points =
    the first 200 FAST(50, nms) corners of the base image.  Returns prev [B,480,640] u8, next, pts [B,200,2] f32, flow [B,2]."""
    g = config2_fixture()
    base, pts0 = g["gray0"], g["pts200"]
    prev = np.repeat(base[None], num_pairs, 0).copy(); nxt = np.zeros_like(prev)
    pts = np.repeat(pts0[None], num_pairs, 0).astype(np.float32).copy(); flow = np.zeros((num_pairs, 2), np.float32)
    for i in range(num_pairs):
        rng = np.random.Generator(np.random.Philox(key=3_000_017 + first_seq + i))
        tx, ty = rng.uniform(-max_shift, max_shift, 2); sh = rng.uniform(-max_shear, max_shear)
        M = np.array([[1.0, sh, tx - sh * 240.0], [0.0, 1.0, ty]])
        nxt[i] = _warp_affine_reflect101(base, M)
        flow[i] = (tx, ty)
    return prev, nxt, pts, flow


def klt_pairs_config2(num_pairs: int):
    """Config 2 itself, tiled to a batch: pair 2k = test -> moved, pair 2k+1 = test -> shear, the 200 FAST corners each.  Returns
    prev, next, pts and the tracked counts cv2 4.13 obtains per pair (190 / 194, status of the golden fixture)."""
    g = config2_fixture()
    prev = np.repeat(g["gray0"][None], num_pairs, 0).copy()
    nxt = np.stack([g["gray_moved"] if i % 2 == 0 else g["gray_shear"] for i in range(num_pairs)])
    pts = np.repeat(g["pts200"][None], num_pairs, 0).astype(np.float32).copy()
    expect = np.array([int(g["moved_200_status"].sum()) if i % 2 == 0 else int(g["shear_200_status"].sum()) for i in range(num_pairs)])
    return prev, nxt, pts, expect


def vio_sequences(first_seq: int, num_seq: int, num_frames: int, w: int = 640, h: int = 480, speed: float = 3.0):
    """frames [T, S, h, w] u8: sequence g = first_seq + i pans over its texture with a constant velocity
    (|v| <= speed px/frame, drawn from Philox(key=g)) — the camera-translation case of the frame loop."""
    m = 64
    assert speed * num_frames < m, "pan leaves the texture margin"
    frames = np.zeros((num_frames, num_seq, h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    textures = {}
    for i in range(num_seq):
        g = first_seq + i
        rng = np.random.Generator(np.random.Philox(key=2_000_003 + g))
        tkey = g % 8
        if tkey not in textures:
            textures[tkey] = texture(np.random.Generator(np.random.Philox(key=7_000_001 + tkey)), h + 2 * m, w + 2 * m)
        vx, vy = rng.uniform(-speed, speed, 2)
        for t in range(num_frames):
            frames[t, i] = np.clip(np.rint(_bilinear(textures[tkey], xx - vx * t + m, yy - vy * t + m)), 0, 255).astype(np.uint8)
    return frames
