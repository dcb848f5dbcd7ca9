"""Times process(dt) alone on the headline batch; flag 0x800 (debug) stops after linearisation + state."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
uv, meas, _ = workload.ekf_streams(0, F, n, 2)
for flags in (0, 0x800):
    b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(np.full(F, n, np.int32), uv)
    for _ in range(3): b.process(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): b.process(0.05)
    e1.record(); torch.cuda.synchronize()
    print(f"flags={flags:#x}: process {e0.elapsed_time(e1) / 20:.4f} ms")
    b.close()
