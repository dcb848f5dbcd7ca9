"""One-step errors of the batch against the FP64 oracle and against the extended-precision evaluation of the same step, per route.
python tools/step_error_probe.py [flags] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
from tests import oracle_lib as O
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
F, n = 128, 50; N = 22 + 3 * n
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); passed = np.ones((F, n), np.uint8)
b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(np.full(F, n, np.int32), uv)
check = list(range(9)) + [23, 100, 127]
so = O.OracleFilter(); so.add_features(uv[0])
dR = torch.from_numpy(R).cuda(); dp = torch.from_numpy(passed).cuda(); dm = torch.from_numpy(meas).cuda()
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
rows = []
before = b.get_state()
for s in range(steps):
    b.process(0.05); b.update(dm[s], dR, dp)
    after = b.get_state_range(0, F)
    for f in check:
        so.set_state(mu=before["mu"][f], feat=before["feat"][f, :n], Pm=before["P"][f, :N, :N], cache=before["cache"][f])
        so.process(0.05); mid = so.state(); so.update(meas[s, f], R[f], passed[f]); os_ = so.state()
        ex = O.step_extended(before["mu"][f], before["feat"][f, :n], before["P"][f, :N, :N], before["cache"][f], 0.05, meas[s, f], R[f], passed[f])
        of = np.r_[os_["mu"], os_["feat"].ravel()]; xf = np.r_[ex["mu"], ex["feat"].ravel()]; gf = np.r_[after["mu"][f], after["feat"][f].ravel()]
        gP = after["P"][f, :N, :N]
        idx = np.array([[22 + 3 * i, 23 + 3 * i] for i in range(n)]).ravel()
        S = mid["P"][np.ix_(idx, idx)] + 1e-5 * np.eye(2 * n)
        ev = np.linalg.eigvalsh((S + S.T) / 2)
        rows.append(dict(s=s, f=f, route=int(after["route"][f]), e=max(rel(gf, of), rel(gP, os_["P"])), e_orc=max(rel(of, xf), rel(os_["P"], ex["P"])),
                         e_true=max(rel(gf, xf), rel(gP, ex["P"])), condS=ev[-1] / max(ev[0], 1e-300) if ev[0] > 0 else -1.0, pmax=float(np.abs(mid["P"]).max()), amax=float(np.abs(os_["P"]).max())))
    before = after
rows.sort(key=lambda r: -r["e"] / max(1e-9, 20 * r["e_orc"]))
for r in rows[:25]:
    print("step %(s)3d filter %(f)3d route %(route)d  e %(e).2e  e_orc %(e_orc).2e  e_true %(e_true).2e  cond(S) %(condS).2e  max|P| prior %(pmax).2e post %(amax).2e" % r)
for rt in (0, 1):
    rr = [r for r in rows if r["route"] == rt]
    if rr:
        print(f"route {rt}: {len(rr)} steps, e>1e-9: {sum(r['e'] > 1e-9 for r in rr)}, failing gate: {sum(r['e'] > max(1e-9, 20 * r['e_orc']) for r in rr)}, max e {max(r['e'] for r in rr):.2e}, max e_orc {max(r['e_orc'] for r in rr):.2e}")
