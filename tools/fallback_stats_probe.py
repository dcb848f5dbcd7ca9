"""How many filters of the config-3 batch ekf_update_fused leaves to the tiled kernels per step, and how sticky that is:
per step the count, how many of them were also left alone in the previous step, and how many are new."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = 50
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
prev = np.zeros(F, bool); ever = np.zeros(F, bool)
rows = []
for s in range(steps):
    b.process(0.05); b.update(dm[s], R, ps)
    route = b.get_state_range(0, F, want_P=False)["route"]
    fb = route != 3
    rows.append((s, int(fb.sum()), int((fb & prev).sum()), int((fb & ~prev).sum()), int((fb & ~ever).sum())))
    prev = fb; ever |= fb
print("step count sticky new first-time")
for r in rows:
    if r[1] or r[0] % 10 == 0: print(*r)
cnt = np.array([r[1] for r in rows]); new = np.array([r[3] for r in rows])
print(f"steps with any fallback filter: {(cnt > 0).sum()} / {steps}; with a new one: {(new > 0).sum()}; mean count {cnt.mean():.2f}; ever: {int(ever.sum())}")
