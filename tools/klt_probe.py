"""Small KLT driver for profiling: pyramid builds + track on synthetic pairs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prev, nxt, pts, flow = workload.klt_pairs(0, B, 640, 480, 200)
trk = capi.KltTracker(640, 480, B, 200)
dp = torch.from_numpy(prev).cuda(); dn = torch.from_numpy(nxt).cuda(); dpts = torch.from_numpy(pts).cuda()
out = dpts.clone(); st = torch.zeros(B, 200, dtype=torch.uint8, device="cuda"); er = torch.zeros(B, 200, device="cuda"); npts = torch.full((B,), 200, dtype=torch.int32, device="cuda")
trk.enable_timing(True)
for it in range(4):
    trk.build_pyramid(0, dp, True); trk.build_pyramid(1, dn, False)
    out.copy_(dpts); trk.track(0, 1, dpts, out, st, er, npts)
torch.cuda.synchronize()
ms, cnt = trk.timing()
print("per-launch ms:", (ms / np.maximum(cnt, 1)).round(4), "tracked", int(st.sum()), "of", B * 200)
# separate timings of the two slots' builds and of the pair build (CUDA events on torch's current stream)
def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("build prev (copy + Scharr + pyrDown) ms:", round(timed(lambda: trk.build_pyramid(0, dp, True)), 4),
      " build next (copy + pyrDown) ms:", round(timed(lambda: trk.build_pyramid(1, dn, False)), 4),
      " pair build ms:", round(timed(lambda: trk.build_pyramid_pair(0, dp, 1, dn)), 4))
