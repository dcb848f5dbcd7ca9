"""Generates tests/golden/ekf_golden.npz from the FP64 oracle (oracle/ekf_oracle.hpp): regression
vectors for the scenarios the reference only prints (test/analyzeEKFSimulation.cpp:233-244,
test/test_ekf.cpp:66-82,156-207, test/jacobian_test.cpp:34-47).  The reference records no expected
numbers for these, so the vectors pin the oracle against drift, not against the
reference; the two reference-held known answers are asserted in tests/test_oracle_ekf.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import oracle_lib as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ekf_golden.npz")


def main():
    out = {}
    for sid, sc in enumerate(O.SCENARIOS[:5]):
        steps, uv, meas = O.scenario(**sc)
        o = O.OracleFilter(); o.add_features(uv)
        dt = float(np.float32(sc["dt"])); n = sc["n"]
        R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (n, 1)); ps = np.ones(n, np.uint8)
        for s in range(steps):
            o.process(dt); o.update(meas[s].astype(np.float64), R, ps)
        st = o.state()
        out[f"s{sid}_steps"] = np.array(steps); out[f"s{sid}_uv0"] = uv[:4]; out[f"s{sid}_meas_last"] = meas[-1, :4]
        out[f"s{sid}_mu"] = st["mu"]; out[f"s{sid}_feat"] = st["feat"]; out[f"s{sid}_Pdiag"] = np.diag(st["P"]); out[f"s{sid}_Psum"] = np.array(st["P"].sum())
    feats = np.array([[0.1, 0.1], [-0.1, -0.1], [0.1, -0.1]])
    o = O.OracleFilter(depth_var=0.0, uv_var=0.0); o.add_features(feats)
    out["jac_dt01"] = o.linearize(0.1)
    mu = o.state()["mu"]; mu[10] = 3.1415; mu[7] = 1.0
    st = o.state(); o.set_state(mu=mu, feat=st["feat"], Pm=st["P"], cache=st["cache"])
    out["jac_omega_v"] = o.linearize(0.1)
    out["jac_stale_dt02"] = o.linearize(0.2)          # E2: columns 7-9 use the dq_inv cached for dt = 0.1
    o2 = O.OracleFilter(); o2.add_features(feats)
    o2.update(feats, np.tile(np.array([1e-3, 0, 0, 1e-3]), (3, 1)), np.array([1, 0, 1], np.uint8))
    s2 = o2.state(); out["upd3_mu"] = s2["mu"]; out["upd3_feat"] = s2["feat"]; out["upd3_P"] = s2["P"]
    g = np.zeros(32, np.float32)
    O.ekf.ekfo_rng_gaussian(O.C.c_uint64(0), 32, O.P(g)); out["rng_gauss32"] = g
    u = np.zeros(8); O.ekf.ekfo_rng_uniform(O.C.c_uint64(0), 8, O.C.c_double(-1.5), O.C.c_double(1.5), O.P(u)); out["rng_uniform8"] = u
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
