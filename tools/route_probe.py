"""Per-step behaviour of the routed default path: the default batch runs free; before every step the general (LDL^T) batch is
seeded with the default batch's state, both take the step, and the results are compared by the route the default batch took."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = 50; N = 22 + 3 * n
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
d = capi.EkfBatch(F, n); g = capi.EkfBatch(F, n, params=capi.default_params(capi.FLAG_FORCE_GENERAL_PATH))
for x in (d, g):
    x.add_features_h(np.full(F, n, np.int32), uv)
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
worst = {0: 0.0, 1: 0.0, 2: 0.0}; count = {0: 0, 1: 0, 2: 0}
shown = 0
for s in range(steps):
    sd = d.get_state()
    g.set_state(mu=sd["mu"], feat=sd["feat"], P=sd["P"], cache=sd["cache"], flags=sd["flags"], klt_last=sd["klt_last"])
    for x in (d, g):
        x.process(0.05); x.update(dm[s], R, ps)
    a = d.get_state_range(0, F); b = g.get_state()
    for f in range(F):
        e = max(rel(a["mu"][f], b["mu"][f]), rel(a["P"][f], b["P"][f]))
        r = int(a["route"][f]); count[r] += 1
        worst[r] = max(worst[r], e) if np.isfinite(e) else float("inf")
        if (r != 0 or e > 1e-8) and shown < 60:
            shown += 1
            ev = np.linalg.eigvalsh((sd["P"][f][:N, :N] + sd["P"][f][:N, :N].T) / 2)
            print(f"step {s} filter {f} route {r} status {a['status'][f]} err {e:.3e} prior: max|P| {np.abs(sd['P'][f]).max():.3e} mineig {ev[0]:.3e} mindiag {np.diag(sd['P'][f])[:N].min():.3e}; after max|P| {np.abs(b['P'][f]).max():.3e}")
print("worst per route", worst, "count", count)
