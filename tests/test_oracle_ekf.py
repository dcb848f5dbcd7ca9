"""CPU tests of the EKF oracle: the reference's two known answers, the cv::RNG restatement, the
simulation harness, and the committed regression vectors.  No GPU."""
import os

import numpy as np
import pytest

from tests import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ekf_golden.npz"))


def test_known_answer_measurement_map():
    """test/test_ekf.cpp:51-63: H for measured {T,F,T} has ones at (0,22),(1,23),(2,28),(3,29)."""
    o = O.OracleFilter()
    o.add_features(np.array([[0.1, 0.1], [-0.1, -0.1], [0.1, -0.1]]))
    H = o.measurement_map(np.array([1, 0, 1], np.uint8))
    ans = np.zeros((4, 31)); ans[0, 22] = ans[1, 23] = ans[2, 28] = ans[3, 29] = 1.0
    np.testing.assert_array_equal(H, ans)


def test_known_answer_resize_keeps_block():
    """test/test_ekf.cpp:27-37 (conservativeResize keeps the top-left block) as addNewFeatures uses it."""
    o = O.OracleFilter()
    o.add_features(np.array([[0.1, 0.2]]))
    before = o.state()["P"].copy()
    before[0, 1] = 2.0; before[1, 0] = 3.0
    st = o.state(); o.set_state(mu=st["mu"], feat=st["feat"], Pm=before)
    o.add_features(np.array([[0.3, 0.4], [0.5, 0.6]]))
    after = o.state()["P"]
    assert after.shape == (31, 31)
    np.testing.assert_array_equal(after[:25, :25], before)
    assert np.count_nonzero(after[25:, :25]) == 0 and np.count_nonzero(after[:25, 25:]) == 0
    np.testing.assert_allclose(np.diag(after)[25:], [1e-5, 1e-5, 100, 1e-5, 1e-5, 100])


def test_initial_state_and_feature_init():
    """initializeBaseState (TightlyCoupledEKF.cpp:23-56) and Feature(homogenous, depth) (Feature.cpp:14-20)."""
    o = O.OracleFilter()
    s = o.state()
    assert s["mu"][3] == 1.0 and np.count_nonzero(s["mu"]) == 1
    np.testing.assert_array_equal(np.diag(s["P"]), [0] * 7 + [30] * 9 + [0.5] * 6)
    o.add_features(np.array([[0.25, -0.5]]))
    s = o.state()
    np.testing.assert_array_equal(s["feat"][0], [0.25, -0.5, 2.0])
    np.testing.assert_array_equal(s["klt_last"][0], [0.25, -0.5])


def test_cv_rng_gaussian_stream():
    g = np.zeros(32, np.float32)
    O.ekf.ekfo_rng_gaussian(O.C.c_uint64(0), 32, O.P(g))
    np.testing.assert_array_equal(g, GOLD["rng_gauss32"])
    cv2 = pytest.importorskip("cv2")
    n = 4000
    g = np.zeros(n, np.float32)
    O.ekf.ekfo_rng_gaussian(O.C.c_uint64(0), n, O.P(g))
    cv2.setRNGSeed(0)
    a = np.zeros((n, 1), np.float32)
    cv2.randn(a, 0, 1)
    np.testing.assert_array_equal(g, a[:, 0])             # bit-exact against OpenCV's own generator


def test_scenario_step_counts_and_landmarks():
    """`for(float t = dt; t <= tf; t += dt)` in float gives 9 and 99 steps (analyzeEKFSimulation.cpp:45)."""
    counts = [O.scenario(**sc)[0] for sc in O.SCENARIOS]
    assert counts == [9, 99, 99, 99, 99, 99]
    steps, uv, meas = O.scenario(**O.SCENARIOS[0])
    assert np.all(np.abs(uv) <= 1.5 + 1e-6) and meas.shape == (9, 30, 2)
    # pure x translation at 0.5 m/s, depth 0.5: u moves by -0.05 per step
    np.testing.assert_allclose(meas[0, :, 0] - uv[:, 0], -0.05, atol=2e-6)
    np.testing.assert_allclose(meas[0, :, 1], uv[:, 1], atol=1e-6)


@pytest.mark.parametrize("sid", range(5))
def test_simulation_matches_golden(sid):
    sc = O.SCENARIOS[sid]
    steps, uv, meas = O.scenario(**sc)
    assert steps == int(GOLD[f"s{sid}_steps"])
    np.testing.assert_array_equal(uv[:4], GOLD[f"s{sid}_uv0"])
    o = O.OracleFilter(); o.add_features(uv)
    dt = float(np.float32(sc["dt"])); n = sc["n"]
    R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (n, 1)); ps = np.ones(n, np.uint8)
    for s in range(steps):
        o.process(dt)
        neg, asym = o.check_sigma()
        assert neg == 0 and asym <= 1e-3                 # the reference's pass criterion (checkSigma)
        o.update(meas[s].astype(np.float64), R, ps)
    st = o.state()
    np.testing.assert_allclose(st["mu"], GOLD[f"s{sid}_mu"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st["feat"], GOLD[f"s{sid}_feat"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.diag(st["P"]), GOLD[f"s{sid}_Pdiag"], rtol=1e-8, atol=1e-14)
    if sid == 0:                                          # 9 steps at 0.5 m/s: the filter has converged to p_x = 0.225
        assert abs(st["mu"][0] - 0.225) < 1e-5 and abs(st["mu"][7] - 0.5) < 1e-4
    assert np.linalg.eigvalsh((st["P"] + st["P"].T) / 2).min() > 0


def test_jacobian_structure_and_stale_cache():
    feats = np.array([[0.1, 0.1], [-0.1, -0.1], [0.1, -0.1]])
    o = O.OracleFilter(depth_var=0.0, uv_var=0.0); o.add_features(feats)
    F = o.linearize(0.1)
    np.testing.assert_allclose(F, GOLD["jac_dt01"], rtol=1e-9, atol=1e-12)
    # F = [[A, 0], [B, D]]: features do not depend on p, q or the biases; D is block diagonal
    assert np.count_nonzero(F[22:, :7]) == 0 and np.count_nonzero(F[22:, 16:22]) == 0 and np.count_nonzero(F[:22, 22:]) == 0
    np.testing.assert_array_equal(F[16:22, 16:22], np.eye(6))
    for i in range(3):
        for j in range(3):
            if i != j:
                assert np.count_nonzero(F[22 + 3 * i:25 + 3 * i, 22 + 3 * j:25 + 3 * j]) == 0
    mu = o.state()["mu"]; mu[10] = 3.1415; mu[7] = 1.0
    st = o.state(); o.set_state(mu=mu, feat=st["feat"], Pm=st["P"], cache=st["cache"])
    np.testing.assert_allclose(o.linearize(0.1), GOLD["jac_omega_v"], rtol=1e-9, atol=1e-12)
    stale = o.linearize(0.2)
    np.testing.assert_allclose(stale, GOLD["jac_stale_dt02"], rtol=1e-9, atol=1e-12)
    # E2 is visible: a filter that never saw dt = 0.1 gives different columns 7..9
    o2 = O.OracleFilter(depth_var=0.0, uv_var=0.0); o2.add_features(feats)
    o2.set_state(mu=mu, feat=st["feat"], Pm=st["P"])
    fresh = o2.linearize(0.2)
    assert np.abs(fresh[22:, 7:10] - stale[22:, 7:10]).max() > 1e-3
    assert np.abs(fresh[:, 10:] - stale[:, 10:]).max() < 1e-12


def test_update_golden_and_flags():
    feats = np.array([[0.1, 0.1], [-0.1, -0.1], [0.1, -0.1]])
    o = O.OracleFilter(); o.add_features(feats)
    o.update(feats, np.tile(np.array([1e-3, 0, 0, 1e-3]), (3, 1)), np.array([1, 0, 1], np.uint8))
    s = o.state()
    np.testing.assert_allclose(s["P"], GOLD["upd3_P"], rtol=1e-9, atol=1e-15)
    assert list(s["flags"]) == [0, 1, 0]
    # measured uv variance: P = 1e-5, R = 1e-3  ->  P' = P R / (P + R)
    np.testing.assert_allclose(s["P"][22, 22], 1e-5 * 1e-3 / (1e-5 + 1e-3), rtol=1e-9)
    assert s["P"][25, 25] == 1e-5 and s["P"][24, 24] == 100.0


def test_float_instantiation_tracks_double_loosely():
    """The literal float32 instantiation agrees with the FP64 oracle only to ~1e-2."""
    sc = O.SCENARIOS[0]
    steps, uv, meas = O.scenario(**sc)
    res = []
    for use_float in (False, True):
        o = O.OracleFilter(use_float=use_float); o.add_features(uv)
        R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (30, 1)); ps = np.ones(30, np.uint8)
        for s in range(steps):
            o.process(float(np.float32(0.05))); o.update(meas[s].astype(np.float64), R, ps)
        res.append(o.state())
    assert np.abs(res[0]["mu"] - res[1]["mu"]).max() < 5e-2
    assert np.abs(res[0]["mu"][0] - res[1]["mu"][0]) > 1e-9      # and they are genuinely different computations
