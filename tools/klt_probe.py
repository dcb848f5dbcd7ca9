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
