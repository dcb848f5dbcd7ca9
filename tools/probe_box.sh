#!/bin/bash
# First-contact probe of the GPU box: host cores, GPU, FP64 peaks (own microbench + cuBLAS DGEMM via torch).
mkdir -p gpurun_out
{
echo "nproc=$(nproc)"; lscpu | grep -E "Model name|Socket|Core|Thread" ; free -g | head -2
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv
./tools/fp64_peak
python - <<'PY'
import torch, time, json
r = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    r[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
for (B, M, K) in ((4096, 172, 100), (4096, 172, 200), (64, 922, 600)):
    a = torch.randn(B, M, K, dtype=torch.float64, device="cuda"); b = torch.randn(B, K, M, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.bmm(a, b)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); torch.bmm(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    r[f"bmm_{B}x{M}x{K}_tflops"] = 2 * B * M * M * K / best * 1e-9
print(json.dumps(r))
PY
} 2>&1 | tee gpurun_out/probe_box.log
