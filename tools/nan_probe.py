import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ekf_vio_b200 import capi, workload
F, n, steps = 4096, 50, 203
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(np.full(F, n, np.int32), uv)
mu = np.zeros((F, 22)); ft = np.zeros((F, n, 3))
for s in range(steps):
    b.process(0.05); b.update(dm[s], R, ps)
    if s % 10 == 9 or s == steps - 1:
        b.read_mu_h(mu, ft)
        bad = np.nonzero(~np.isfinite(mu).all(1))[0]
        if len(bad):
            st = b.get_state(want_P=False)
            print("step", s + 1, "non-finite filters", bad[:10], "status", st["status"][bad[:10]], "mu", mu[bad[0]][:8])
            break
else:
    print("all finite after", steps, "steps; flags", flags)
