"""The bench's KLT step (both pyramids by reference + track + postprocess) a few times, for ncu.  python tools/klt_step_probe.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
npts = 200
prev, nxt, pts, _ = workload.klt_pairs_8d(0, 32)
t = lambda a: np.ascontiguousarray(np.concatenate([a] * ((B + 31) // 32))[:B])
prev, nxt, pts = t(prev), t(nxt), t(pts)
trk = capi.KltTracker(640, 480, B, npts)
dp, dn, dpts = torch.from_numpy(prev).cuda(), torch.from_numpy(nxt).cuda(), torch.from_numpy(pts).cuda()
out = torch.zeros_like(dpts); st = torch.zeros(B, npts, dtype=torch.uint8, device="cuda"); er = torch.zeros(B, npts, device="cuda")
n = torch.full((B,), npts, dtype=torch.int32, device="cuda")
for _ in range(steps):
    trk.build_pyramid_pair_ref(0, dp, 1, dn, False)
    out.copy_(dpts); trk.track(0, 1, dpts, out, st, er, n)
torch.cuda.synchronize()
print("tracked", float(st.float().mean()))
