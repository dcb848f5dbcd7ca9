import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch, ctypes
from ekf_vio_b200 import capi, workload
F, n = 4096, 50
uv, meas, _ = workload.ekf_streams(0, F, n, 2)
nf = np.full(F, n, np.int32)
names = ["pro issue", "pro wait", "C+base", "grp wait", "U/fb/scale", "frag+bar", "dmma+store"]
for flags in (0, 0x2000, 0x4000, 0x6000):
    b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(nf, uv)
    for _ in range(3): b.process(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): b.process(0.05)
    e1.record(); torch.cuda.synchronize()
    print(f"flags={flags:#x}: process {e0.elapsed_time(e1) / 20:.4f} ms", flush=True)
    if hasattr(capi.lib, "ekfvio_debug_cov_clocks"):
        buf = (ctypes.c_ulonglong * 16)()
        capi.lib.ekfvio_debug_cov_clocks(buf, 1)
        b.process(0.05); torch.cuda.synchronize()
        capi.lib.ekfvio_debug_cov_clocks(buf, 1)
        print("   " + "  ".join(f"{nm}={buf[i] / F:.0f}" for i, nm in enumerate(names)))
    b.close()
