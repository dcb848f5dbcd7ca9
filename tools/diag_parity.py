"""Diagnostic: free-running divergence of each GPU path from the oracle, per scenario."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import oracle_lib as O
from ekf_vio_b200 import capi

def rel(a, b): return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)

def run(sid, flags, steps_max=40):
    sc = O.SCENARIOS[sid]; steps, uv, meas = O.scenario(**sc); n = sc["n"]
    orc = O.OracleFilter(); orc.add_features(uv)
    b = capi.EkfBatch(1, n, params=capi.default_params(flags))
    b.add_features_h(np.array([n], np.int32), uv.astype(np.float64)[None])
    dt = float(np.float32(sc["dt"])); R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (1, n, 1)); ps = np.ones((1, n), np.uint8)
    dR = torch.from_numpy(R).cuda(); dp = torch.from_numpy(ps).cuda()
    out = []
    for s in range(min(steps, steps_max)):
        orc.process(dt); b.process(dt)
        z = meas[s].astype(np.float64)
        orc.update(z, R[0], ps[0]); b.update(torch.from_numpy(z[None].copy()).cuda(), dR, dp)
        g = b.get_state(); o = orc.state()
        N = 22 + 3 * n
        out.append((rel(np.concatenate([g["mu"][0], g["feat"][0].ravel()]), np.concatenate([o["mu"], o["feat"].ravel()])), rel(g["P"][0, :N, :N], o["P"])))
    return out

for sid in (0, 1, 4, 5):
    for name, fl in (("general", 1), ("tiled", 0), ("gen-gain+tiled-joseph", 0x100), ("tiled-gain+gen-joseph", 0x200)):
        o = run(sid, fl)
        mu = max(x[0] for x in o); P = max(x[1] for x in o)
        print(f"scenario {sid} {name:24s} worst mu {mu:.2e} worst P {P:.2e}   first steps P: " + " ".join(f"{x[1]:.1e}" for x in o[:8]))
