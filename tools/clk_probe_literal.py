import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from ekf_vio_b200 import capi, workload
F, n = 592, 50
uv, meas, _ = workload.ekf_streams(0, F, n, 4)
b = capi.EkfBatch(F, n, params=capi.default_params(capi.FLAG_LITERAL_JOSEPH)); b.add_features_h(np.full(F, n, np.int32), uv)
R = torch.from_numpy(np.tile(np.array([1e-5,0,0,1e-5]), (F,n,1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
out = (C.c_ulonglong * 8)()
b.enable_timing(True)
for s in range(4):
    b.process(0.05); b.update(dm[s], R, ps); torch.cuda.synchronize()
    capi.lib.ekfvio_debug_clocks(out, 1)
    v = np.array(list(out), dtype=np.float64) / F
    print("step", s, "joseph marks", v[:4].round(0), "solve marks [setup, S+fwd, bwd+store, W+store]", v[4:8].round(0))
kms, kcnt = b.timing()
print({k: round(float(kms[i] / max(kcnt[i], 1)), 4) for i, k in enumerate(["process", "chol", "cov", "fwd/solve", "fused"])})
