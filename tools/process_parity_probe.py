"""One process(dt) of the batch against the FP64 oracle seeded with the batch's own state before it (config-3 streams, after a few
process + update steps): max-norm relative error of Sigma' (lower form mirrored by get_state) and of the state, per kernel variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
from tests import oracle_lib as O
F, n, steps = 64, 50, 6
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); passed = np.ones((F, n), np.uint8)
dR = torch.from_numpy(R).cuda(); dp = torch.from_numpy(passed).cuda()
for flags in (0, 0x1000):
    b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(np.full(F, n, np.int32), uv)
    for s in range(steps):
        b.process(0.05); b.update(torch.from_numpy(meas[s]).cuda(), dR, dp)
    s0 = b.get_state()
    b.process(0.05)
    s1 = b.get_state()
    worst_P = worst_mu = 0.0
    for f in range(0, F, 4):
        o = O.OracleFilter(); o.add_features(uv[f])
        o.set_state(mu=s0["mu"][f], feat=s0["feat"][f, :n], Pm=s0["P"][f][:22 + 3 * n, :22 + 3 * n], cache=s0["cache"][f])
        o.process(0.05)
        so = o.state()
        N = 22 + 3 * n
        worst_P = max(worst_P, np.abs(s1["P"][f][:N, :N] - so["P"]).max() / np.abs(so["P"]).max())
        worst_mu = max(worst_mu, np.abs(s1["mu"][f] - so["mu"]).max() / np.abs(so["mu"]).max())
    print(f"flags={flags:#x}: one process step vs oracle: Sigma {worst_P:.3e}  mu {worst_mu:.3e}")
    b.close()
