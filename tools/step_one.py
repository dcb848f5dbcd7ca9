"""A few process + update steps of the config-3 batch (4096 filters x 50 features) for ncu captures of the EKF kernels:
ncu --set full --clock-control none --import-source on -k regex:"ekf_update_fused|ekf_process_cov_tiles" -s 4 -c 2 python tools/step_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F, n, steps = 4096, 50, 4
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda()
ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
for s in range(steps):
    b.process(0.05)
    b.update(dm[s], R, ps)
torch.cuda.synchronize()
