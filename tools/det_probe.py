import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ekf_vio_b200 import capi, workload
F, n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 50, int(sys.argv[2]) if len(sys.argv) > 2 else 12
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
outs = []
for rep in range(3):
    b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
    for s in range(steps):
        b.process(0.05); b.update(dm[s], R, ps)
    st = b.get_state(); outs.append(st); b.close()
for k in ("mu", "feat", "P"):
    print(k, [bool(np.array_equal(outs[0][k], outs[i][k])) for i in (1, 2)], np.abs(outs[0][k] - outs[1][k]).max())
# host-buffer arm
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
Rh = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); ph = np.ones((F, n), np.uint8); mu = np.zeros((F, 22)); ft = np.zeros((F, n, 3))
for s in range(steps):
    b.process(0.05); b.update_h(meas[s], Rh, ph); b.read_mu_h(mu, ft)
st = b.get_state()
print("host arm equal:", bool(np.array_equal(st["mu"], outs[0]["mu"])), bool(np.array_equal(st["P"], outs[0]["P"])), np.abs(st["P"] - outs[0]["P"]).max())
