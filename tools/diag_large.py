import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ekf_vio_b200 import capi, workload
from tests import oracle_lib as O
n, F, steps = 80, 4, 40
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); ps = np.ones((F, n), np.uint8)
dR = torch.from_numpy(R).cuda(); dps = torch.from_numpy(ps).cuda(); dm = torch.from_numpy(meas).cuda()
bs = {"default": capi.EkfBatch(F, n, params=capi.default_params(0)), "literal": capi.EkfBatch(F, n, params=capi.default_params(4))}
for b in bs.values(): b.add_features_h(np.full(F, n, np.int32), uv)
orc = O.OracleFilter(); orc.add_features(uv[0])
rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
for s in range(steps):
    orc.process(0.05); orc.update(meas[s, 0], R[0], ps[0])
    for b in bs.values(): b.process(0.05); b.update(dm[s], dR, dps)
    if s % 5 == 4 or s < 3:
        o = orc.state()
        out = []
        for k, b in bs.items():
            st = b.get_state()
            out.append(f"{k}: P {rel(st['P'][0], o['P']):.2e} mu {rel(st['mu'][0], o['mu']):.2e}")
        st0, st1 = bs['default'].get_state(), bs['literal'].get_state()
        print(s + 1, " | ".join(out), f"| default vs literal P {rel(st0['P'], st1['P']):.2e}")
