import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F, n = 4096, 50
uv, meas, _ = workload.ekf_streams(0, F, n, 2)
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
for _ in range(3): b.process(0.05)
torch.cuda.synchronize()
