"""Track kernel time on the bench's KLT workload.  python tools/track_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
B, npts = 256, 200
for name, gen in (("8d", lambda: workload.klt_pairs_8d(0, 32)), ("r1-synthetic", lambda: workload.klt_pairs(0, 32, 640, 480, 200))):
    prev, nxt, pts, _ = gen()
    t = lambda a: np.ascontiguousarray(np.concatenate([a] * 8)[:B])
    prev, nxt, pts = t(prev), t(nxt), t(pts)
    trk = capi.KltTracker(640, 480, B, npts)
    dp, dn, dpts = torch.from_numpy(prev).cuda(), torch.from_numpy(nxt).cuda(), torch.from_numpy(pts).cuda()
    out = torch.zeros_like(dpts); st = torch.zeros(B, npts, dtype=torch.uint8, device="cuda"); er = torch.zeros(B, npts, device="cuda")
    n = torch.full((B,), npts, dtype=torch.int32, device="cuda")
    trk.build_pyramid_pair(0, dp, 1, dn, False)
    for _ in range(3):
        out.copy_(dpts); trk.track(0, 1, dpts, out, st, er, n)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out.copy_(dpts); trk.track(0, 1, dpts, out, st, er, n)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: stride mode {os.environ.get('EKFVIO_KLT_TRACK_STRIDE', '1')}: track {e0.elapsed_time(e1) / 10:.3f} ms, tracked {float(st.float().mean()):.3f}")
    trk.close()
