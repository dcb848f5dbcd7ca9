"""ekf_update_fused probe: (1) run-to-run determinism, (2) fused vs the three-kernel update, (3) per-kernel event timing,
(4) phase clocks when EKFVIO_LIB_PATH points at the `make prof` build.  Usage: python tools/fused_probe.py [F] [steps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ekf_vio_b200 import capi, workload  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n = 50
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda()
ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
has_clk = hasattr(capi.lib, "ekfvio_debug_fused_clocks")


def run(tag, timing=False):
    b = capi.EkfBatch(F, n)
    b.add_features_h(np.full(F, n, np.int32), uv)
    if timing:
        b.enable_timing(True)
    routes = np.zeros(5, np.int64)
    for s in range(steps):
        b.process(0.05)
        b.update(dm[s], R, ps)
        if not timing:
            st = b.get_state_range(0, F, want_P=False)
            routes += np.bincount(st["route"] + 1, minlength=5)
    torch.cuda.synchronize()
    if timing:
        kms, kcnt = b.timing()
        print(tag, "kernel ms:", {k: round(float(kms[i] / max(kcnt[i], 1)), 4) for i, k in enumerate(["process", "chol", "cov", "fwd/solve", "fused"])})
        if has_clk:
            out = (C.c_ulonglong * 24)()
            capi.lib.ekfvio_debug_fused_clocks(out, 1)
            v = np.array(list(out), dtype=np.float64) / (F * steps * 3)        # three runs since the last reset
            names = ["map", "load", "panel", "update", "barrier", "store", "tail", "-"]
            print(tag, "clocks per filter, warp 0 :", dict(zip(names, v[:8].round(0))), "sum", v[:8].sum().round(0))
            print(tag, "clocks per filter, warp 15:", dict(zip(names, v[8:16].round(0))), "sum", v[8:16].sum().round(0))
            print(tag, "factor_block per filter [load, LDL, checks+rs, inverse, store]:", v[16:21].round(0), "tail: own copy-out done / barrier passed", v[21:23].round(0))
    else:
        print(tag, "routes [-1, sym, joseph_sym, joseph_full, done]:", routes.tolist())
    st = b.get_state()
    b.close()
    return st


a = run("fused run 1")
b = run("fused run 2")
print("deterministic: mu", np.array_equal(a["mu"], b["mu"], equal_nan=True), "P", np.array_equal(a["P"], b["P"], equal_nan=True),
      "feat", np.array_equal(a["feat"], b["feat"], equal_nan=True))
if not np.array_equal(a["P"], b["P"], equal_nan=True):
    d = np.abs(a["P"] - b["P"]).reshape(F, -1).max(1)
    print("  filters that differ:", np.nonzero(d > 0)[0][:20], "max", d.max())
run("fused timed", timing=True)
if has_clk:
    pass
os.environ["EKFVIO_NO_FUSED_UPDATE"] = "1"
c = run("three-kernel")
c2 = run("three-kernel run 2")
print("three-kernel deterministic: P", np.array_equal(c["P"], c2["P"], equal_nan=True))
run("three-kernel timed", timing=True)
den = np.abs(c["P"]).reshape(F, -1).max(1)
rel = np.abs(a["P"] - c["P"]).reshape(F, -1).max(1) / den
print("fused vs three-kernel after", steps, "steps: rel max per filter: median %.2e  p99 %.2e  max %.2e" % (np.median(rel), np.quantile(rel, 0.99), rel.max()))
print("sym of fused P:", np.abs(a["P"] - a["P"].transpose(0, 2, 1)).max())
