"""Config 4 probe: 64 filters x 300 features (N = 922, m = 600); times process / gain / covariance."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = 4
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
b = capi.EkfBatch(F, n); b.add_features_h(np.full(F, n, np.int32), uv)
R = torch.from_numpy(np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))).cuda(); ps = torch.ones(F, n, dtype=torch.uint8, device="cuda")
dm = torch.from_numpy(meas).cuda()
b.process(0.05); b.update(dm[0], R, ps); torch.cuda.synchronize()
b.enable_timing(True)
t0 = time.perf_counter()
for s in range(1, steps):
    b.process(0.05); b.update(dm[s], R, ps)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / (steps - 1)
ms, cnt = b.timing()
st = b.get_state(want_P=False)
N, m = 22 + 3 * n, 2 * n
fl = 486 * n * n + 5.7e3 * n + 4.3e4 + 1.75e3 * n + 5.3e3 + 4 * N * N * m + 2 * N * m * m + m ** 3 / 3
print(f"F={F} n={n}: {dt*1e3:.2f} ms/step, {F/dt:.0f} filter-steps/s, {F/dt*fl/1e12:.2f} TFLOP/s algorithmic; per-kernel ms", (ms / np.maximum(cnt, 1)).round(3)[:4], "status", st["status"].max(), "finite", np.isfinite(st["mu"]).all())
