"""Which filters of the config-3 stream (v, omega ~ U(-0.2, 0.2)) leave the well-conditioned regime, on which
update path, and what the FP64 oracle does on the same filters.  Run on the GPU box: python tools/stability_probe.py [F] [steps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
from tests import oracle_lib as O
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = 50
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
out = {}
first_bad = {}
for name, flags in (("default", 0), ("literal", 4), ("general", 1)):
    b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(np.full(F, n, np.int32), uv)
    neg = torch.zeros(F, dtype=torch.int32, device="cuda"); asym = torch.zeros(F, dtype=torch.float64, device="cuda")
    fb = np.full(F, -1)
    maxP = np.zeros((steps, F))
    for s in range(steps):
        b.process(0.05); b.update(dm[s], R, ps)
        st = b.get_state(want_P=(name == "default" and F <= 1024))
        bad = st["status"] != 0
        fb[(fb < 0) & bad] = s
        if st["P"] is not None:
            maxP[s] = np.abs(st["P"]).reshape(F, -1).max(1)
    b.check_sigma(neg, asym); torch.cuda.synchronize()
    out[name] = dict(status_nonzero=int((st["status"] != 0).sum()), neg_diag_filters=int((neg > 0).sum().item()), max_asym=float(asym[torch.isfinite(asym)].max().item()) if bool(torch.isfinite(asym).any()) else None,
                     nonfinite=int((~np.isfinite(st["mu"]).all(1)).sum()), first_bad_hist=np.bincount(fb[fb >= 0] // 10, minlength=steps // 10).tolist())
    first_bad[name] = fb
    if name == "default":
        mp = maxP
    b.close()
print(json.dumps(out, indent=1))
badf = np.where(first_bad["default"] >= 0)[0][:6]
print("default-path flagged filters:", badf.tolist(), "first flagged at", first_bad["default"][badf].tolist())
Rn = np.tile(np.array([1e-5, 0, 0, 1e-5]), (n, 1)); pn = np.ones(n, np.uint8)
for f in badf[:3]:
    o = O.OracleFilter(); o.add_features(uv[f])
    tr = []
    for s in range(steps):
        o.process(0.05); o.update(meas[s, f], Rn, pn)
        P = o.state()["P"]
        tr.append((float(np.abs(P).max()), float(np.diag(P).min()), int(o.state()["status"][0])))
    s0 = first_bad["default"][f]
    print(f"filter {f}: flagged at {s0}; oracle max|P| around: ", [f"{t[0]:.2e}" for t in tr[max(0, s0 - 4):s0 + 2]], " gpu max|P|:", [f"{x:.2e}" for x in mp[max(0, s0 - 4):s0 + 2, f]],
          "oracle min diag overall %.2e" % min(t[1] for t in tr), "oracle status any", any(t[2] for t in tr), "oracle max|P| overall %.2e" % max(t[0] for t in tr))
