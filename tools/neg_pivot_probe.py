"""Debug: a step with an S that is not positive definite, taken from the free-running default batch, replayed on 1-filter batches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
from tests import oracle_lib as O
F, n = 128, 50; N = 22 + 3 * n
steps = 100
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); passed = np.ones((F, n), np.uint8)
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
dR = torch.from_numpy(R).cuda(); dp = torch.from_numpy(passed).cuda(); dm = torch.from_numpy(meas).cuda()
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
found = 0
for s in range(steps):
    before = b.get_state()
    b.process(0.05); b.update(dm[s], dR, dp)
    after = b.get_state_range(0, F)
    for f in np.where((after["route"] == 1))[0]:
        if found >= 6: break
        P0 = before["P"][f]; P0 = (P0 + P0.T) / 2
        o = O.OracleFilter(); o.add_features(uv[f]); o.set_state(mu=before["mu"][f], feat=before["feat"][f], Pm=P0[:N, :N], cache=before["cache"][f])
        o.process(0.05); o.update(meas[s, f], R[f], passed[f]); os_ = o.state()
        out = [f"step {s} filter {f} status {after['status'][f]} oracle status {os_['status'][0]} batch-vs-oracle {rel(after['P'][f][:N,:N], os_['P']):.2e}"]
        for name, fl in (("default", 0), ("literal", 4), ("general", 1), ("default-nolower(0x400)", 0x400)):
            x = capi.EkfBatch(1, n, params=capi.default_params(fl)); x.add_features_h(np.array([n], np.int32), uv[f:f + 1])
            x.set_state(mu=before["mu"][f:f + 1], feat=before["feat"][f:f + 1], P=P0[None], cache=before["cache"][f:f + 1])
            x.process(0.05); x.update(dm[s][f:f + 1].contiguous(), dR[f:f + 1].contiguous(), dp[f:f + 1].contiguous())
            a = x.get_state_range(0, 1)
            out.append(f"{name}: route {a['route'][0]} err {max(rel(a['P'][0][:N,:N], os_['P']), rel(a['mu'][0], os_['mu'])):.2e}")
            x.close()
        print(" | ".join(out)); found += 1
    if found >= 6: break
