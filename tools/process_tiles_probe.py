"""process(dt) on DMMA tiles (ekf_process_cov_tiles) against the row-block kernel (debug flag 0x1000): same batch, same steps —
lower form of Sigma compared element by element, then both timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ekf_vio_b200 import capi, workload
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
steps = 4
uv, meas, _ = workload.ekf_streams(0, F, n, steps)
nf = np.full(F, n, np.int32)
if len(sys.argv) > 3:   # ragged feature counts
    nf = (np.arange(F) % (n + 1)).astype(np.int32)
R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); passed = np.ones((F, n), np.uint8)
dR = torch.from_numpy(R).cuda(); dp = torch.from_numpy(passed).cuda()
out = {}
for flags in (0, 0x1000):
    b = capi.EkfBatch(F, n, params=capi.default_params(flags)); b.add_features_h(nf, uv)
    for s in range(steps):
        b.process(0.05)
        b.update(torch.from_numpy(meas[s]).cuda(), dR, dp)
    b.process(0.05)
    out[flags] = b.get_state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): b.process(0.05)
    e0.record()
    for _ in range(20): b.process(0.05)
    e1.record(); torch.cuda.synchronize()
    print(f"flags={flags:#x}: process {e0.elapsed_time(e1) / 20:.4f} ms  launches={b.launches}", flush=True)
    b.close()
a, c = out[0], out[0x1000]
worst = 0.0
for f in range(F):
    N = 22 + 3 * int(nf[f])
    pa, pc = a["P"][f][:N, :N], c["P"][f][:N, :N]
    d = np.abs(pa - pc).max() / max(np.abs(pc).max(), 1e-300)
    worst = max(worst, d)
    if d > 1e-12 and f < 4096:
        i, j = np.unravel_index(np.abs(pa - pc).argmax(), pa.shape)
        print(f"filter {f} n={nf[f]}: rel {d:.3e} at ({i},{j}): {pa[i, j]} vs {pc[i, j]}"); 
        if f > 8: break
print("worst rel diff of Sigma:", worst, " mu:", np.abs(a["mu"] - c["mu"]).max(), " sym:", max(np.abs(a["P"][f] - a["P"][f].T).max() for f in range(min(F, 64))))
if hasattr(capi.lib, "ekfvio_debug_cov_clocks"):
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    b = capi.EkfBatch(F, n); b.add_features_h(nf, uv)
    b.process(0.05); torch.cuda.synchronize()
    capi.lib.ekfvio_debug_cov_clocks(buf, 1)
    b.process(0.05); torch.cuda.synchronize()
    capi.lib.ekfvio_debug_cov_clocks(buf, 1)
    names = ["pro issue", "pro wait", "C+base", "grp wait", "blockscale", "T gemm", "fb gemm", "frag+bar", "dmma+store"]
    tot = sum(buf[i] for i in range(9))
    for i, nm in enumerate(names): print(f"  {nm:12s} {buf[i] / F:9.0f} clk/filter  {100.0 * buf[i] / tot:5.1f}%")
    print(f"  total {tot / F:.0f} clk/filter")
