"""Times the replenishFeatures path (FAST + NMS + compaction, greedy scan) on a batch of 640x480 frames."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prev, nxt, pts, flow = workload.klt_pairs(0, B, 640, 480, 200)
det = capi.FastDetector(640, 480, B, 4096)
d = torch.from_numpy(prev).cuda()
kp = torch.zeros(B, 4096, 2, dtype=torch.int16, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
needed = torch.full((B,), 100, dtype=torch.int32, device="cuda")
new_px = torch.zeros(B, 128, 2, dtype=torch.int16, device="cuda"); n_new = torch.zeros(B, dtype=torch.int32, device="cuda")
ex = torch.from_numpy(pts[:, :60].copy()).cuda(); nex = torch.full((B,), 60, dtype=torch.int32, device="cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_det = timed(lambda: det.detect(d, 50, True, kp, None, cnt))
t_sel = timed(lambda: det.select(kp, cnt, ex, nex, needed, 30, 11, None, new_px, None, n_new))
print(f"B={B}: detect {t_det:.3f} ms ({B / t_det * 1e3:.0f} frames/s, {B * 640 * 480 / t_det / 1e6:.1f} Gpx/s), select {t_sel:.3f} ms; "
      f"keypoints/frame {cnt.float().mean().item():.0f}, new/frame {n_new.float().mean().item():.1f}")
try:
    import cv2
    fd = cv2.FastFeatureDetector_create(threshold=50, nonmaxSuppression=True)
    t0 = time.perf_counter()
    for i in range(16): fd.detect(prev[i])
    print(f"cv2 FAST on this host: {(time.perf_counter() - t0) / 16 * 1e3:.3f} ms/frame (1 call at a time, {cv2.getNumThreads()} threads)")
except ImportError:
    pass
