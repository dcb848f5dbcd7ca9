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
