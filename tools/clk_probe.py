import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from ekf_vio_b200 import capi, workload
F, n = 1184, 50
uv, meas, _ = workload.ekf_streams(0, F, n, 6)
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
R = torch.from_numpy(np.tile(np.array([1e-5,0,0,1e-5]), (F,n,1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
out = (C.c_ulonglong * 8)()
for s in range(6):
    b.process(0.05); b.update(dm[s], R, ps); torch.cuda.synchronize()
    capi.lib.ekfvio_debug_clocks(out, 1)
    v = np.array(list(out), dtype=np.float64) / F
    print("step", s, "joseph marks", v[:4].round(0), "solve marks [setup, S+fwd, bwd+store, W+store]", v[4:8].round(0))
