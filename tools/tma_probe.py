"""TMA-staged pyramid kernels against the cp.async kernels (EKFVIO_KLT_NO_TMA=1) on the same frames: bit-exactness of every level
and derivative image, and the time of a pair build.  python tools/tma_probe.py [batch] [w] [h]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
prev, nxt, pts, _ = workload.klt_pairs(0, min(B, 8), w, h, 16)
reps = (B + len(prev) - 1) // len(prev)
prev = np.ascontiguousarray(np.concatenate([prev] * reps)[:B]); nxt = np.ascontiguousarray(np.concatenate([nxt] * reps)[:B])
dp, dn = torch.from_numpy(prev).cuda(), torch.from_numpy(nxt).cuda()
os.environ["EKFVIO_KLT_NO_TMA"] = "1"
ref = capi.KltTracker(w, h, B, 16)
del os.environ["EKFVIO_KLT_NO_TMA"]
tma = capi.KltTracker(w, h, B, 16)
for t in (ref, tma):
    t.build_pyramid_pair(0, dp, 1, dn, False)
torch.cuda.synchronize()
print("levels", tma.num_levels, "launches per pair build: cp.async", ref.launches, "tma", tma.launches)
bad = 0
for slot, wd in ((0, True), (1, False)):
    for img in sorted({0, B // 2, B - 1}):
        for lv in range(tma.num_levels):
            a = ref.read_level(slot, img, lv, wd); b = tma.read_level(slot, img, lv, wd)
            for x, y, nm in zip(a, b, ("img", "deriv")):
                if x is None: continue
                if not np.array_equal(x, y):
                    bad += 1
                    d = np.argwhere(np.asarray(x) != np.asarray(y))
                    print(f"MISMATCH slot {slot} img {img} level {lv} {nm}: {len(d)} elements, first {d[:4].tolist()}")
print("bit-exact" if bad == 0 else f"{bad} mismatching planes")
for name, t in (("cp.async", ref), ("tma", tma)):
    for _ in range(3): t.build_pyramid_pair(0, dp, 1, dn, False)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): t.build_pyramid_pair(0, dp, 1, dn, False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms*1e3:.1f} us per pair-batch build, {B * (2140800 + 508800) / (ms * 1e-3) / 1e9:.0f} GB/s algorithmic")
