"""GPU parity tests of the pyramidal KLT tracker through the C ABI.

Bars (north_star / SURVEY.md §8d): pyramid levels and Scharr derivatives bit-exact; tracked/lost
status bit-exact; positions within 0.01 px on points that pass the reference's kill-pad test.
Checked against (a) golden vectors produced by cv2 4.13.0 (tests/golden/make_klt_golden.py),
(b) the C oracle on seeded synthetic pairs, (c) cv2 itself when importable.
"""
import os
import zlib

import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "klt_config2.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def make_tracker(w, h, batch, max_points, **kw):
    from ekf_vio_b200 import capi
    p = capi.default_klt_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return capi.KltTracker(w, h, batch, max_points, params=p)


def track_h(trk, prev, nxt, pts, init=None):
    """prev/nxt [B,H,W] u8, pts [B,n,2] -> next [B,n,2], status [B,n], err [B,n]"""
    B, n = pts.shape[0], pts.shape[1]
    pp = np.zeros((B, trk.max_points, 2), np.float32); pp[:, :n] = pts
    nn = pp.copy() if init is None else np.zeros_like(pp)
    if init is not None:
        nn[:, :n] = init
    st, er = trk.track_pair_h(np.ascontiguousarray(prev), np.ascontiguousarray(nxt), pp, nn, np.full(B, n, np.int32))
    return nn[:, :n], st[:, :n], er[:, :n]


def in_killpad(p, w, h, pad=11):
    return ~((p[..., 0] < pad) | (p[..., 1] < pad) | (w - p[..., 0] < pad) | (h - p[..., 1] < pad))


def test_pyramid_and_scharr_bit_exact_vs_cv2_golden(cuda, gold):
    import torch
    g0 = gold["gray0"]
    trk = make_tracker(640, 480, 2, 8)
    assert trk.num_levels == 4
    imgs = torch.from_numpy(np.stack([g0, gold["gray_moved"]])).cuda()
    trk.build_pyramid(0, imgs, True)
    for l in range(4):
        img, der = trk.read_level(0, 0, l, True)
        np.testing.assert_array_equal(img, gold[f"pyr{l}"])
        assert zlib.crc32(np.ascontiguousarray(der).tobytes()) == int(gold[f"scharr{l}_crc32"])
        if l >= 2:
            np.testing.assert_array_equal(der, gold[f"scharr{l}"])
        # second image of the batch against the oracle
        img1, der1 = trk.read_level(0, 1, l, True)
        oi, od, _ = O.klt_level(gold["gray_moved"], 21, 3, l)
        np.testing.assert_array_equal(img1, oi)
        np.testing.assert_array_equal(der1, od)


@pytest.mark.parametrize("w,h", [(77, 101), (333, 245), (64, 32), (641, 479), (130, 66)])
def test_pyramid_odd_and_small_sizes_vs_oracle(cuda, w, h):
    import torch
    rng = np.random.default_rng(w * 1000 + h)
    img = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    trk = make_tracker(w, h, 3, 4)
    # caller pitch larger than the width, unaligned
    pitch = w + 5
    buf = np.zeros((3, h, pitch), np.uint8); buf[:, :, :w] = img
    trk.build_pyramid(1, torch.from_numpy(buf).cuda(), True)
    _, _, ml = O.klt_level(img[0], 21, 3, 0)
    assert trk.num_levels == ml + 1
    for b in range(3):
        for l in range(ml + 1):
            gi, gd = trk.read_level(1, b, l, True)
            oi, od, _ = O.klt_level(img[b], 21, 3, l)
            np.testing.assert_array_equal(gi, oi)
            np.testing.assert_array_equal(gd, od)


@pytest.mark.parametrize("pair", ["moved", "shear"])
def test_config2_tracking_vs_cv2_golden(cuda, gold, pair):
    """Config 2: 200 FAST corners, test -> moved / shear, the reference's LK arguments."""
    g0, g1 = gold["gray0"], gold[f"gray_{pair}"]
    pts = gold["pts200"]
    trk = make_tracker(640, 480, 1, 256)
    nx, st, er = track_h(trk, g0[None], g1[None], pts[None])
    ref_nx, ref_st = gold[f"{pair}_200_next"], gold[f"{pair}_200_status"]
    np.testing.assert_array_equal(st[0], ref_st)                     # tracked / lost bit-exact
    assert int(st.sum()) == {"moved": 190, "shear": 194}[pair]
    ok = (ref_st == 1) & in_killpad(ref_nx, 640, 480)
    d = np.abs(nx[0][ok] - ref_nx[ok]).max()
    assert d <= 0.01, d
    # and against the oracle the agreement is to float rounding
    on, ost, oer, _ = O.klt_calc_optical_flow(g0, g1, pts, pts)
    np.testing.assert_array_equal(st[0], ost)
    both = ost == 1
    assert np.abs(nx[0][both] - on[both]).max() <= 1e-4
    assert np.abs(er[0][both] - oer[both]).max() <= 1e-3
    print(f"{pair}: max |dpos| vs cv2 {d:.2e}")


@pytest.mark.parametrize("pair", ["moved", "shear"])
def test_stress_points_status_bit_exact(cuda, gold, pair):
    """1366 points incl. texture-less and out-of-image ones: status must equal cv2's on every point."""
    g0, g1 = gold["gray0"], gold[f"gray_{pair}"]
    pts = gold["pts_stress"]
    trk = make_tracker(640, 480, 1, 1408)
    nx, st, _ = track_h(trk, g0[None], g1[None], pts[None])
    ref_nx, ref_st = gold[f"{pair}_stress_next"], gold[f"{pair}_stress_status"]
    np.testing.assert_array_equal(st[0], ref_st)
    ok = (ref_st == 1) & in_killpad(ref_nx, 640, 480) & in_killpad(nx[0], 640, 480)
    # gate on kill-pad passers; the FAST corners (first 666) must all be within 0.01 px
    d_fast = np.abs(nx[0][:666][ok[:666]] - ref_nx[:666][ok[:666]]).max()
    assert d_fast <= 0.01, d_fast
    on, ost, _, _ = O.klt_calc_optical_flow(g0, g1, pts, pts)
    np.testing.assert_array_equal(st[0], ost)
    both = ost == 1
    assert np.abs(nx[0][both] - on[both]).max() <= 2e-3


def synth_pair(rng, w, h, shift):
    """Smooth random texture and a translated copy (integer shift keeps it exact)."""
    big = rng.integers(0, 256, (h // 4 + 16, w // 4 + 16)).astype(np.float32)
    big = np.kron(big, np.ones((4, 4), np.float32))
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16
    for _ in range(2):
        big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, big)
        big = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 0, big)
    big = big.astype(np.uint8)
    a = big[16:16 + h, 16:16 + w]
    b = big[16 - shift[1]:16 - shift[1] + h, 16 - shift[0]:16 - shift[0] + w]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def test_batched_synthetic_pairs_vs_oracle(cuda):
    """A batch of independent image pairs with ragged point counts; each must match the oracle."""
    rng = np.random.default_rng(3)
    B, w, h, mp = 5, 320, 240, 96
    prevs, nexts, ptss, counts = [], [], [], [96, 1, 0, 50, 77]
    for b in range(B):
        a, c = synth_pair(rng, w, h, (int(rng.integers(-6, 7)), int(rng.integers(-6, 7))))
        prevs.append(a); nexts.append(c)
        ptss.append(np.stack([rng.uniform(-3, w + 3, mp), rng.uniform(-3, h + 3, mp)], 1).astype(np.float32))
    trk = make_tracker(w, h, B, mp)
    pp = np.stack(ptss); nn = pp.copy()
    st, er = trk.track_pair_h(np.stack(prevs), np.stack(nexts), pp, nn, np.array(counts, np.int32))
    for b in range(B):
        c = counts[b]
        on, ost, oer, _ = O.klt_calc_optical_flow(prevs[b], nexts[b], ptss[b][:c], ptss[b][:c])
        np.testing.assert_array_equal(st[b, :c], ost)
        both = ost == 1
        if both.any():
            assert np.abs(nn[b, :c][both] - on[both]).max() <= 2e-3
        assert not st[b, c:].any()                  # entries beyond npts are untouched


@pytest.mark.parametrize("win", [7, 9, 15, 27, 31])
def test_other_window_sizes_vs_oracle(cuda, win):
    """Window sizes around the tracker's tile-staging cases (aligned words with an offset where the row stride has room for it,
    the byte loop at 31): status as the oracle's on points inside, on and beyond every border."""
    rng = np.random.default_rng(100 + win)
    w, h, mp = 320, 240, 128
    a, c = synth_pair(rng, w, h, (int(rng.integers(-4, 5)), int(rng.integers(-4, 5))))
    pts = np.stack([rng.uniform(-3, w + 3, mp), rng.uniform(-3, h + 3, mp)], 1).astype(np.float32)
    pts[:8] = [[0, 0], [w - 1, h - 1], [win / 2, win / 2], [w - win / 2, 5], [1.5, h / 2], [w / 2, 1.5], [w - 1.5, h / 2], [w / 2, h - 1.5]]
    trk = make_tracker(w, h, 1, mp, window_size=win)
    nn = pts[None].copy()
    st, er = trk.track_pair_h(a[None], c[None], pts[None], nn, np.array([mp], np.int32))
    on, ost, oer, _ = O.klt_calc_optical_flow(a, c, pts, pts, win=win)
    np.testing.assert_array_equal(st[0], ost)
    both = ost == 1
    assert both.sum() > mp // 2
    assert np.abs(nn[0][both] - on[both]).max() <= 2e-3


def test_property_known_translation_at_full_size(cuda):
    """Size-independent property at BASELINE's 640x480: a pure integer translation is recovered."""
    rng = np.random.default_rng(9)
    a, c = synth_pair(rng, 640, 480, (5, -3))
    pts = np.stack([rng.uniform(60, 580, 200), rng.uniform(60, 420, 200)], 1).astype(np.float32)
    trk = make_tracker(640, 480, 1, 200)
    nx, st, _ = track_h(trk, a[None], c[None], pts[None])
    good = st[0] == 1
    assert good.mean() > 0.9
    flow = nx[0][good] - pts[good]
    assert np.abs(np.median(flow, 0) - np.array([5, -3])).max() < 0.05


def test_postprocess_matches_reference_epilogue(cuda, gold):
    """KLTTracker.cpp:72-92 incl. the K linear-index error E1 (principal point dropped)."""
    import torch
    nx, st = gold["moved_200_next"], gold["moved_200_status"]
    K = np.zeros((3, 3), np.float32); K[0, 0] = 400.0; K[1, 1] = 410.0; K[0, 2] = 320.0; K[1, 2] = 240.0; K[2, 2] = 1
    K9 = np.ascontiguousarray(K.T).reshape(9)        # column-major storage, as Eigen's Matrix3f
    trk = make_tracker(640, 480, 1, 200)
    dn = torch.from_numpy(nx[None].copy()).cuda(); ds = torch.from_numpy(st[None].copy()).cuda()
    npts = torch.tensor([200], dtype=torch.int32, device="cuda")
    meas = torch.full((1, 200, 2), -7.0, dtype=torch.float32, device="cuda"); cov = torch.zeros(1, 200, 4, dtype=torch.float32, device="cuda")
    passed = torch.zeros(1, 200, dtype=torch.uint8, device="cuda")
    trk.postprocess(dn, ds, npts, torch.from_numpy(K9[None].copy()).cuda(), meas, cov, passed)
    om, oc, op = O.klt_postprocess(nx, st, 640, 480, K9)
    np.testing.assert_array_equal(passed[0].cpu().numpy(), op)
    m = meas[0].cpu().numpy()
    np.testing.assert_array_equal(m[op == 1], om[op == 1])
    assert (m[op == 0] == -7.0).all()                # failed entries untouched, as in the reference
    np.testing.assert_array_equal(cov[0].cpu().numpy(), oc)
    assert abs(m[op == 1][0, 0] * 400.0 - nx[op == 1][0, 0]) < 1e-3   # E1: no principal point subtracted


def test_cv2_live_if_available(cuda, gold):
    cv2 = pytest.importorskip("cv2")
    g0, g1 = gold["gray0"], gold["gray_shear"]
    pts = gold["pts200"]
    init = pts + np.float32(1.5)                      # a non-trivial initial flow
    p0 = pts.reshape(-1, 1, 2).copy(); p1 = init.reshape(-1, 1, 2).copy()
    rn, rs, _ = cv2.calcOpticalFlowPyrLK(g0, g1, p0, p1, winSize=(21, 21), maxLevel=3,
                                         criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW, minEigThreshold=1e-4)
    trk = make_tracker(640, 480, 1, 200)
    nx, st, _ = track_h(trk, g0[None], g1[None], pts[None], init[None])
    np.testing.assert_array_equal(st[0], rs[:, 0])
    ok = (rs[:, 0] == 1) & in_killpad(rn[:, 0], 640, 480)
    assert np.abs(nx[0][ok] - rn[:, 0][ok]).max() <= 0.01


def test_track_next_h_equals_track_pair_h_along_a_sequence(cuda):
    """ekfvio_klt_track_next_h (only the new frame uploaded, previous pyramid reused) must give bit-identical positions and status
    to ekfvio_klt_track_pair_h on the same consecutive frames (KLTTracker as EKFVIO::addFrame drives it, EKFVIO.cpp:201-217)."""
    from ekf_vio_b200 import capi, workload
    T, B, npts = 5, 3, 64
    frames = workload.vio_sequences(0, B, T, 320, 240, speed=2.5)           # [T, B, h, w]
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(30, [290, 210], (npts, 2)) for _ in range(B)]).astype(np.float32)
    n = np.array([npts, npts - 7, 20], np.int32)
    a = capi.KltTracker(320, 240, B, npts); b = capi.KltTracker(320, 240, B, npts)
    for t in range(1, T):
        na = pts.copy(); nb = pts.copy()
        sa, ea = a.track_pair_h(frames[t - 1], frames[t], pts, na, n)
        if t == 1:
            sb, eb = b.track_pair_h(frames[0], frames[1], pts, nb, n)
        else:
            sb, eb = b.track_next_h(frames[t], pts, nb, n)
        for i in range(B):
            np.testing.assert_array_equal(sa[i, :n[i]], sb[i, :n[i]], err_msg=f"frame {t} image {i} status")
            np.testing.assert_array_equal(na[i, :n[i]], nb[i, :n[i]], err_msg=f"frame {t} image {i} positions")
            np.testing.assert_array_equal(ea[i, :n[i]], eb[i, :n[i]])
    a.close(); b.close()


def test_stats_allreduce_single_rank(cuda):
    """ekfvio_stats_allreduce on a one-rank communicator is the identity (the multi-rank path runs in bench.py --gpus N)."""
    import torch
    from ekf_vio_b200 import capi
    comm = capi.StatsComm(0, 1, 0, lambda raw: raw)
    assert comm.size() == 1
    x = torch.arange(8, dtype=torch.float64, device="cuda")
    comm.allreduce(x); torch.cuda.synchronize()
    np.testing.assert_array_equal(x.cpu().numpy(), np.arange(8.0))
    comm.close()


def test_pyramid_pair_by_reference_tracks_like_the_copying_build(cuda):
    """ekfvio_klt_build_pyramid_pair_ref keeps level 0 in the caller's buffers (no pass-through copy): levels, derivatives, tracked
    positions and status must be bit-identical to the copying build — aligned and unaligned pitch (TMA and cp.async staging)."""
    import torch
    from ekf_vio_b200 import capi, workload
    for w, h, pitch in ((640, 480, 640), (320, 240, 328), (333, 197, 333)):
        B, npts = 3, 48
        prev, nxt, pts, _ = workload.klt_pairs(5, B, w, h, npts)
        pp = np.zeros((B, h, pitch), np.uint8); nn_ = np.zeros_like(pp); pp[:, :, :w] = prev; nn_[:, :, :w] = nxt
        dp, dn = torch.from_numpy(pp).cuda(), torch.from_numpy(nn_).cuda()
        dpts = torch.from_numpy(pts).cuda(); n = torch.full((B,), npts, dtype=torch.int32, device="cuda")
        res = []
        for by_ref in (False, True):
            t = capi.KltTracker(w, h, B, npts)
            (t.build_pyramid_pair_ref if by_ref else t.build_pyramid_pair)(0, dp, 1, dn, False)
            out = dpts.clone(); st = torch.zeros(B, npts, dtype=torch.uint8, device="cuda"); er = torch.zeros(B, npts, device="cuda")
            t.track(0, 1, dpts, out, st, er, n); torch.cuda.synchronize()
            lv = [t.read_level(0, 1, l, True) for l in range(t.num_levels)]
            res.append((out.cpu().numpy(), st.cpu().numpy(), er.cpu().numpy(), lv))
            t.close()
        np.testing.assert_array_equal(res[0][0], res[1][0]); np.testing.assert_array_equal(res[0][1], res[1][1]); np.testing.assert_array_equal(res[0][2], res[1][2])
        for a, b in zip(res[0][3], res[1][3]):
            np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])
        assert res[0][1].sum() > 0.8 * B * npts


def test_sample_based_uncertainty_matches_getrectsubpix_restatement(cuda):
    """KLTTracker::estimateUncertaintySampleBased (KLTTracker.cpp:111-175, dead code in the reference: parity unpinned there).
    Oracle: the reference's loop restated with OpenCV's own cv2.getRectSubPix(patchType=CV_32F) — interior features, features
    whose samples reach over the image border (replicated), integer and fractional positions."""
    cv2 = pytest.importorskip("cv2")
    from ekf_vio_b200 import capi, workload
    g = workload.config2_fixture()
    ref_img, cur_img = g["gray0"], g["gray_moved"]
    rng = np.random.default_rng(9)
    mu_ref = np.concatenate([g["pts200"][:40], rng.uniform([2, 2], [637, 477], (20, 2)), np.array([[3.0, 3.0], [636.5, 476.25], [320.0, 5.0]])]).astype(np.float32)
    mu = (mu_ref + rng.uniform(-3, 3, mu_ref.shape)).astype(np.float32)
    mu[:10] = np.round(mu[:10])                                   # integer positions: a = max(0, 1e-4) branch of getRectSubPix
    got = capi.klt_sample_uncertainty_h(ref_img, cur_img, mu_ref, mu)
    want = np.zeros_like(got)
    for i in range(len(mu)):
        ref = cv2.getRectSubPix(ref_img, (5, 5), (float(mu_ref[i, 0]), float(mu_ref[i, 1])), patchType=cv2.CV_32F)
        s = sxx = syy = sxy = np.float32(0)
        for du in (-10.0, -5.0, 0.0, 5.0, 10.0):
            for dv in (-10.0, -5.0, 0.0, 5.0, 10.0):
                smp = cv2.getRectSubPix(cur_img, (5, 5), (float(np.float32(mu[i, 0] + np.float32(du))), float(np.float32(mu[i, 1] + np.float32(dv)))), patchType=cv2.CV_32F)
                ssd = np.float32(0)
                for a in range(5):
                    for b in range(5):
                        ssd = np.float32(ssd + np.float32(np.float64(ref[a, b] - smp[a, b]) ** 2))
                ssd = np.float32(ssd / np.float32(25))
                rd = np.float32(np.exp(np.float64(np.float32(-0.01) * ssd)))
                s = np.float32(s + rd); sxx = np.float32(sxx + np.float32(np.float32(rd * np.float32(du)) * np.float32(du)))
                syy = np.float32(syy + np.float32(np.float32(rd * np.float32(dv)) * np.float32(dv))); sxy = np.float32(sxy + np.float32(np.float32(rd * np.float32(du)) * np.float32(dv)))
        want[i] = [[sxx / s, sxy / s], [sxy / s, syy / s]]
    assert np.isfinite(got).all()
    # float exp / pow of two libraries and OpenCV's vectorised getRectSubPix: agreement to ~1e-4 of each feature's covariance scale
    err = np.abs(got - want).reshape(len(mu), -1).max(1) / np.abs(want).reshape(len(mu), -1).max(1)
    print(f"sample-based uncertainty: worst relative difference {err.max():.2e} over {len(mu)} features")
    assert err.max() <= 5e-4
    assert np.abs(got[:, 0, 1] - got[:, 1, 0]).max() == 0
