"""GPU tests of the device-side frame loop (EKFVIO::addFrame, EKFVIO.cpp:139-196).

The loop is a composition of components that each have their own parity tests against the oracles
(test_gpu_ekf / test_gpu_klt / test_gpu_replenish).  Here the composition itself is checked: the same
frames are pushed (a) through ekfvio_vio_add_frame and (b) through the individual C-ABI calls with the
reference's host-side glue restated in numpy float32 (KLTTracker.cpp:53-59, EKFVIO.cpp:201-217, :246-308);
the filters must end bit-identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def host_composed(frames, K9, dts, num_features, thr=50, min_dist=30, kill_pad=11):
    """The frame loop driven from the host through the component entry points."""
    import torch
    from ekf_vio_b200 import capi
    T, S, h, w = frames.shape
    batch = capi.EkfBatch(S, num_features)
    trk = capi.KltTracker(w, h, S, num_features)
    det = capi.FastDetector(w, h, S, 4096)
    dK = torch.from_numpy(K9).cuda()
    cur = 0
    Kprev = K9.copy()
    for t in range(T):
        d = torch.from_numpy(frames[t]).cuda()
        new_slot = 0 if t == 0 else cur ^ 1
        trk.build_pyramid(new_slot, d, True)
        if t > 0:
            st = batch.get_state(want_P=False)                      # klt_last / nfeat: untouched by process()
            batch.process(torch.from_numpy(dts[t]).cuda())
            # (state only: no need to make the batch complete Sigma between process() and update())
            h_mu = np.zeros((S, 22)); h_feat = np.zeros((S, num_features, 3))
            batch.read_mu_h(h_mu, h_feat)
            n = st["nfeat"]
            prev_pts = np.zeros((S, num_features, 2), np.float32); next_pts = np.zeros_like(prev_pts)
            kl = st["klt_last"].astype(np.float32); ft = h_feat.astype(np.float32)
            for s in range(S):
                Kp, Kc = Kprev[s], K9[s]
                prev_pts[s, :n[s], 0] = kl[s, :n[s], 0] * Kp[0] + Kp[2]; prev_pts[s, :n[s], 1] = kl[s, :n[s], 1] * Kp[4] + Kp[5]
                next_pts[s, :n[s], 0] = Kc[0] * ft[s, :n[s], 0] + Kc[2]; next_pts[s, :n[s], 1] = Kc[4] * ft[s, :n[s], 1] + Kc[5]
            dp, dn = torch.from_numpy(prev_pts).cuda(), torch.from_numpy(next_pts).cuda()
            dst = torch.zeros(S, num_features, dtype=torch.uint8, device="cuda"); der = torch.zeros(S, num_features, device="cuda")
            dnp = torch.from_numpy(n.astype(np.int32)).cuda()
            trk.track(cur, new_slot, dp, dn, dst, der, dnp)
            meas = torch.zeros(S, num_features, 2, device="cuda"); cov = torch.zeros(S, num_features, 4, device="cuda")
            psd = torch.zeros(S, num_features, dtype=torch.uint8, device="cuda")
            trk.postprocess(dn, dst, dnp, dK, meas, cov, psd)
            ok = psd.cpu().numpy().astype(bool) & (np.arange(num_features)[None, :] < n[:, None])
            z = np.where(ok[..., None], meas.cpu().numpy().astype(np.float64), 0.0)
            R = np.where(ok[..., None], cov.cpu().numpy().astype(np.float64), 0.0)
            batch.update(torch.from_numpy(z).cuda(), torch.from_numpy(R).cuda(), torch.from_numpy(ok.astype(np.uint8)).cuda())
        st = batch.get_state(want_P=False)
        n = st["nfeat"]; ft = st["feat"].astype(np.float32)
        ex = np.zeros((S, num_features, 2), np.float32)
        for s in range(S):
            ex[s, :n[s], 0] = K9[s, 0] * ft[s, :n[s], 0] + K9[s, 2]; ex[s, :n[s], 1] = K9[s, 4] * ft[s, :n[s], 1] + K9[s, 5]
        needed = np.maximum(num_features - n, 0).astype(np.int32)
        kp = torch.zeros(S, 4096, 2, dtype=torch.int16, device="cuda"); cnt = torch.zeros(S, dtype=torch.int32, device="cuda")
        det.detect(d, thr, True, kp, None, cnt)
        new_px = torch.zeros(S, num_features, 2, dtype=torch.int16, device="cuda"); new_m = torch.zeros(S, num_features, 2, device="cuda")
        n_new = torch.zeros(S, dtype=torch.int32, device="cuda")
        det.select(kp, cnt, torch.from_numpy(ex).cuda(), torch.from_numpy(n.astype(np.int32)).cuda(), torch.from_numpy(needed).cuda(), min_dist, kill_pad,
                   dK, new_px, new_m, n_new)
        k = np.minimum(n_new.cpu().numpy(), num_features - n).astype(np.int32)
        batch.add_features_h(k, new_m.cpu().numpy().astype(np.float64))
        Kprev = K9.copy()
        cur = new_slot
    out = batch.get_state()
    batch.close(); trk.close(); det.close()
    return out


@pytest.mark.parametrize("graph", [True, False], ids=["cuda-graph", "eager"])
def test_frame_loop_equals_host_composition_and_tracks(cuda, graph):
    """T = 8 frames: eager first frames, one graph capture per pyramid-slot parity (frames 3, 4), replays (5-7)."""
    import torch
    from ekf_vio_b200 import capi, workload
    S, T, w, h, NF = 3, 8, 320, 240, 40
    frames = workload.vio_sequences(0, S, T, w, h, speed=2.0)
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 200.0; K9[:, 4] = 200.0; K9[:, 6] = 160.0; K9[:, 7] = 120.0; K9[:, 8] = 1.0   # column-major
    dts = np.full((T, S), 0.05)
    ref = host_composed(frames, K9, dts, NF)

    loop = capi.VioLoop(S, w, h, num_features=NF, use_cuda_graph=graph)
    dK = torch.from_numpy(K9).cuda()
    for t in range(T):
        loop.add_frame(torch.from_numpy(frames[t]).cuda(), dK, None if t == 0 else torch.from_numpy(dts[t]).cuda())
    torch.cuda.synchronize()
    assert loop.frame_count == T and loop.launches > 0
    got = loop.filters.get_state()
    for key in ("nfeat", "mu", "feat", "klt_last", "P", "status", "flags"):
        np.testing.assert_array_equal(got[key], ref[key], err_msg=key)
    # the loop did something: features were added on the first frame, tracked and fused afterwards
    assert (got["nfeat"] > 10).all() and (got["status"] == 0).all() and np.isfinite(got["mu"]).all()
    moved = np.abs(got["klt_last"][:, :5] - got["feat"][:, :5, :2]).max()
    assert moved < 0.05            # the fused feature positions stay next to the last KLT results
    loop.close()


def test_first_frame_only_replenishes(cuda):
    import torch
    from ekf_vio_b200 import capi, workload
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import replenish_oracle as R
    S, w, h, NF = 2, 320, 240, 30
    frames = workload.vio_sequences(5, S, 1, w, h)
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 250.0; K9[:, 4] = 240.0; K9[:, 8] = 1.0
    loop = capi.VioLoop(S, w, h, num_features=NF)
    loop.add_frame(torch.from_numpy(frames[0]).cuda(), torch.from_numpy(K9).cuda())
    st = loop.filters.get_state(want_P=False)
    for s in range(S):
        kp, _ = R.fast9_16(frames[0, s], 50, True)
        px, metric = R.select_new_features(kp, [], w, h, NF, K9=K9[s])
        assert st["nfeat"][s] == len(px)
        np.testing.assert_array_equal(st["feat"][s, :len(px), :2], metric.astype(np.float64))   # (u, v) of the new features
        np.testing.assert_array_equal(st["klt_last"][s, :len(px)], metric.astype(np.float64))
        assert (st["feat"][s, :len(px), 2] == 2.0).all()                                        # inverse of DEFAULT_POINT_DEPTH 0.5
    assert (st["mu"][:, 3] == 1.0).all()
    loop.close()


def test_frame_loop_with_feature_removal_keeps_tracking(cuda):
    """remove_lost_features = 1 (not in the reference): features whose track is lost leave the state and replenishment
    refills it; without it (the reference's behaviour) lost features pile up and the state only ever grows."""
    import torch
    from ekf_vio_b200 import capi, workload
    S, T, w, h, NF = 2, 9, 320, 240, 30
    frames = workload.vio_sequences(3, S, T, w, h, speed=6.0)           # fast pan: features leave through the kill-pad
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 200.0; K9[:, 4] = 200.0; K9[:, 8] = 1.0
    dK = torch.from_numpy(K9).cuda(); ddt = torch.full((S,), 0.05, dtype=torch.float64, device="cuda")
    out = {}
    for rm in (False, True):
        loop = capi.VioLoop(S, w, h, num_features=NF, remove_lost_features=rm)
        for t in range(T):
            loop.add_frame(torch.from_numpy(frames[t]).cuda(), dK, None if t == 0 else ddt)
        out[rm] = loop.filters.get_state(want_P=False)
        loop.close()
    # graph replay and eager execution must agree bit for bit with removal on as well (T = 9 and T = 10: an odd and an even
    # number of replays — removal makes a frame flip the Sigma buffers an odd number of times)
    for T2 in (9, 10):
        frames2 = workload.vio_sequences(3, S, T2, w, h, speed=6.0)
        res = []
        for graph in (True, False):
            loop = capi.VioLoop(S, w, h, num_features=NF, remove_lost_features=True, use_cuda_graph=graph)
            for t in range(T2):
                loop.add_frame(torch.from_numpy(frames2[t]).cuda(), dK, None if t == 0 else ddt)
            res.append(loop.filters.get_state())
            loop.close()
        for key in ("nfeat", "mu", "feat", "P", "klt_last"):
            np.testing.assert_array_equal(res[0][key], res[1][key], err_msg=f"{key} T={T2}")
    keep, drop = out[False], out[True]
    assert keep["flags"].any(), "the pan was meant to lose some features"
    assert not drop["flags"][:, :].any() or (drop["flags"].sum() < keep["flags"].sum())
    assert (drop["status"] == 0).all() and np.isfinite(drop["mu"]).all()
    assert (drop["nfeat"] <= NF).all() and (drop["nfeat"] > 0).all()


def oracle_composed_sequence(frames, K9, dts, num_features, thr=50, min_dist=30, kill_pad=11):
    """EKFVIO::addFrame / updateStateWithNewImage / replenishFeatures (EKFVIO.cpp:139-196, 201-217, 224-311) driven for ONE sequence
    with nothing but the oracles: the FP64 EKF oracle (oracle/ekf_oracle.hpp), OpenCV's own cv2.calcOpticalFlowPyrLK with the
    reference's arguments (KLTTracker.cpp:61-64) and the numpy FAST / greedy-scan oracle (pinned to cv2).  The host glue is the
    reference's float32 arithmetic.  Yields the oracle filter's state after every frame."""
    import cv2
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import replenish_oracle as RO
    from tests import oracle_lib as O
    T, h, w = frames.shape
    K = K9.astype(np.float32)                      # column-major: K(0) = fx, K(4) = fy, K(2) = K(5) = 0 (E1), K(6), K(7) = principal point
    o = O.OracleFilter()
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    for t in range(T):
        img = frames[t]
        if t > 0:
            o.process(float(dts[t]))                                                             # EKFVIO.cpp:163
            st = o.state(); n = len(st["feat"])
            if n:
                kl = st["klt_last"].astype(np.float32); ft = st["feat"].astype(np.float32)
                prev_pts = np.stack([kl[:, 0] * K[0] + K[2], kl[:, 1] * K[4] + K[5]], 1).astype(np.float32)      # Feature::metric2Pixel (KLTTracker.cpp:53-55)
                init_pts = np.stack([K[0] * ft[:, 0] + K[2], K[4] * ft[:, 1] + K[5]], 1).astype(np.float32)      # Feature::getPixel (:57-59)
                nxt, status, _ = cv2.calcOpticalFlowPyrLK(frames[t - 1], img, prev_pts.reshape(-1, 1, 2), init_pts.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=3,
                                                          criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW, minEigThreshold=1e-4)
                nxt = nxt.reshape(-1, 2); status = status.reshape(-1)
                x, y = nxt[:, 0], nxt[:, 1]
                pad = np.float32(kill_pad)
                ok = (status == 1) & ~((x < pad) | (y < pad) | (np.float32(w) - x < pad) | (np.float32(h) - y < pad))     # KLTTracker.cpp:73, Frame.cpp:44-55
                sx = np.float32((1.0 / float(K[0])) ** 2); sy = np.float32((1.0 / float(K[4])) ** 2)
                cov = np.zeros((n, 4), np.float32); cov[ok, 0] = np.float32(0.00001) * sx; cov[ok, 3] = np.float32(0.00001) * sy                 # :75-83
                meas = np.zeros((n, 2), np.float32)
                meas[ok, 0] = (x[ok] - K[2]) / K[0]; meas[ok, 1] = (y[ok] - K[5]) / K[4]                                                        # pixel2Metric, E1
                o.update(meas.astype(np.float64), cov.astype(np.float64), ok.astype(np.uint8))                                                   # EKFVIO.cpp:217
        st = o.state(); n = len(st["feat"]); ft = st["feat"].astype(np.float32)
        existing = np.stack([K[0] * ft[:, 0] + K[2], K[4] * ft[:, 1] + K[5]], 1) if n else np.zeros((0, 2), np.float32)
        kps, _ = RO.fast9_16(img, thr, True)                                                      # cv::FAST(img, kp, 50, true), EKFVIO.cpp:242
        px, metric = RO.select_new_features(kps, existing, w, h, max(num_features - n, 0), min_dist, kill_pad, K9=K)
        if len(px):
            o.add_features(metric.astype(np.float64))                                             # EKFVIO.cpp:308
        yield o.state()


def test_frame_loop_against_the_oracle_composition(cuda):
    """SURVEY.md §8f-3 against oracles only (not against a host composition of the same GPU kernels): 2 sequences x 5 frames.
    After every frame: the same number of features (FAST + greedy selection bit-exact), the same lost / tracked flags (tracker
    status + kill-pad bit-exact vs cv2), the last KLT results within 1e-5 (tracker positions agree with OpenCV's to ~2e-4 px,
    i.e. ~1e-6 in metric units, and pass through the reference's float32 glue), Sigma within 1e-4, and the state within 2e-3:
    the Kalman gain of the barely observed velocity / acceleration states is of order 1e2, so a 1e-6 difference in a measurement
    legitimately moves those states by 1e-4."""
    import torch
    from ekf_vio_b200 import capi, workload
    S, T, w, h, NF = 2, 5, 320, 240, 30
    frames = workload.vio_sequences(11, S, T, w, h, speed=2.0)
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 210.0; K9[:, 4] = 205.0; K9[:, 6] = 160.0; K9[:, 7] = 120.0; K9[:, 8] = 1.0
    dts = np.full((T, S), 0.05)
    gens = [oracle_composed_sequence(frames[:, s], K9[s], dts[:, s], NF) for s in range(S)]
    loop = capi.VioLoop(S, w, h, num_features=NF, use_cuda_graph=True)
    dK = torch.from_numpy(K9).cuda()
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0
    worst = np.zeros(3)
    for t in range(T):
        loop.add_frame(torch.from_numpy(frames[t]).cuda(), dK, None if t == 0 else torch.from_numpy(dts[t]).cuda())
        got = loop.filters.get_state()
        for s in range(S):
            o = next(gens[s])
            n = len(o["feat"]); N = 22 + 3 * n
            assert got["nfeat"][s] == n, f"frame {t} sequence {s}: {got['nfeat'][s]} features, oracle {n}"
            np.testing.assert_array_equal(got["flags"][s, :n], o["flags"], err_msg=f"frame {t} sequence {s}: lost / tracked flags")
            e = np.array([rel(got["klt_last"][s, :n], o["klt_last"]), rel(got["P"][s, :N, :N], o["P"]),
                          rel(np.concatenate([got["mu"][s], got["feat"][s, :n].ravel()]), np.concatenate([o["mu"], o["feat"].ravel()]))])
            worst = np.maximum(worst, e)
            assert e[0] <= 1e-5 and e[1] <= 1e-4 and e[2] <= 2e-3, f"frame {t} sequence {s}: klt_last {e[0]:.3e} Sigma {e[1]:.3e} state {e[2]:.3e}"
        assert t == 0 or got["nfeat"].min() > 5
    print(f"frame loop vs oracle composition over {T} frames x {S} sequences: worst relative difference klt_last {worst[0]:.3e}, Sigma {worst[1]:.3e}, state {worst[2]:.3e}")
    loop.close()


def test_graph_is_recaptured_when_the_batch_was_stepped_behind_the_loops_back(cuda):
    """The captured graph bakes in the batch's Sigma ping-pong buffer.  A caller that steps the batch handed out by
    ekfvio_vio_filters() between two frames (here: one extra process()) flips that buffer; the loop must notice and record
    again instead of replaying into the stale buffer — graph and eager loop stay bit-identical."""
    import torch
    from ekf_vio_b200 import capi, workload
    S, T, w, h, NF = 2, 9, 320, 240, 24
    frames = workload.vio_sequences(21, S, T, w, h, speed=2.0)
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 200.0; K9[:, 4] = 200.0; K9[:, 8] = 1.0
    dK = torch.from_numpy(K9).cuda(); ddt = torch.full((S,), 0.05, dtype=torch.float64, device="cuda")
    res = []
    for graph in (True, False):
        loop = capi.VioLoop(S, w, h, num_features=NF, use_cuda_graph=graph)
        for t in range(T):
            loop.add_frame(torch.from_numpy(frames[t]).cuda(), dK, None if t == 0 else ddt)
            if t == 5:
                loop.filters.process(0.01)              # behind the loop's back: one Sigma buffer flip
        res.append(loop.filters.get_state())
        loop.close()
    for key in ("nfeat", "mu", "feat", "P", "klt_last", "flags"):
        np.testing.assert_array_equal(res[0][key], res[1][key], err_msg=key)
    assert np.isfinite(res[0]["mu"]).all()
