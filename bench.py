#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native ekf_vio hot paths.

    python bench.py --gpus 1 --steps K --warmup W            (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                     (the CPU path on the host cores)

Prints ONE JSON line.  Headline metric (BASELINE.json): EKF filter-steps/s, one step =
process(dt) + updateWithFeaturePositions(all features measured) for every filter of the batch
(config "batched EKF: 4096 independent filters, IMU state + 50 features").  The same line carries
the KLT metric (features tracked/s at 640x480) under "klt".  `value` is measured with inputs
resident in HBM; `e2e` goes through the host-buffer entry points of the C ABI with the
host<->device copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FEAT = 50
DT = 0.05


def flops_per_filter_step(n: int) -> float:
    """Algorithmic FLOPs of one filter-step: block-structured F P F' + Q and the
    finite-difference evaluations, plus 4 N^2 m + 2 N m^2 + m^3/3 for the update."""
    N, m = 22 + 3 * n, 2 * n
    f_proc = 486 * n * n + 5.7e3 * n + 4.3e4 + (1.75e3 * n + 5.3e3)
    f_upd = 4 * N * N * m + 2 * N * m * m + m ** 3 / 3
    return f_proc + f_upd


# dram__bytes_read.sum + dram__bytes_write.sum of one ekf_update_fused launch over 4096 filters x 50 features (ncu --set full, this round)
NCU_FUSED_DRAM_BYTES = 561.662e6 + 579.152e6     # profiles/r02f_ncu_ekf_summary.txt (the final kernel; r02c's capture: 562.1 + 577.3)


def flops_cov_update(n: int) -> float:
    """Covariance update as the reference writes it (Joseph form): 4 N^2 m."""
    N, m = 22 + 3 * n, 2 * n
    return 4.0 * N * N * m


def flops_cov_update_executed(n: int) -> float:
    """What the default kernel executes for symmetric filters: Sigma - Z Z' restricted to the lower triangle, N^2 m."""
    N, m = 22 + 3 * n, 2 * n
    return 1.0 * N * N * m


def flops_update_fused_executed(n: int) -> float:
    """What ekf_update_fused executes per filter: m/8 block steps, each the panel Z_j = Sigma(:,j) inv(L_j)' (2 DMMA per
    tile row: N*8*8*2) and the rank-8 update of the padded lower triangle (2 DMMA = 1024 FLOPs per 8x8 tile, T(T+1)/2 tiles,
    T = ceil((N+1)/8) tile rows incl. the innovation row).  Neither the forward substitution (N m^2) nor a separate factorisation of S
    (m^3/3) exists in this form; the 8x8 factorisations on the FP64 CUDA cores are ~1 % and not counted."""
    N, m = 22 + 3 * n, 2 * n
    T = (N + 1 + 7) // 8
    steps = (m + 7) // 8
    return steps * (T * 8 * 8 * 8 * 2 + T * (T + 1) / 2 * 1024.0)


def flops_per_filter_step_executed(n: int) -> float:
    """Same process step; update as ekf_update_fused executes it."""
    f_proc = 486 * n * n + 5.7e3 * n + 4.3e4 + (1.75e3 * n + 5.3e3)
    return f_proc + flops_update_fused_executed(n)


def ekf_config(F: int, n: int, world: int) -> dict:
    """The workload both arms (ours and --impl reference) are measured on: BASELINE.json configs[2]."""
    return {"workload": f"batched EKF: {F} filters/GPU x (22+3*{n})-dim state, process(dt)+update(all {n} measured) per step, dt={DT}; SURVEY 8d config 3 streams "
                        "(body velocity and angular rate ~ U(-0.2, 0.2), depth sigma 0.01, R = 1e-5 I)",
            "l2": "working set per step (two 1.0 GB Sigma buffers + 1.3 GB gain panels at 4096 filters) is larger than the 126 MB L2; no flush needed",
            "filters_total": F * world, "features": n,
            "update_form": "library default (ekf_update_fused): the measurement blocks applied one after the other with Sigma resident in registers — the block-sequential "
                           "form of Sigma - Z Z', equal to the batch update because R is block diagonal; per filter and update, where S is positive definite with a "
                           "pivot ratio <= 1e9, term by term (signed factor S = L J L') otherwise; EKFVIO_FLAG_LITERAL_JOSEPH = always term by term"}


KLT_BYTES_WITH_DERIVS = 2_140_800   # SURVEY.md §8d, 640x480 levels 0-3, read once + write levels 1-3 + int16x2 derivatives
KLT_BYTES_NO_DERIVS = 508_800
# dram__bytes_read.sum + dram__bytes_write.sum per image pair of the two pyramid kernels from the round's ncu --set full capture
# (profiles/r02_ncu_klt_summary.txt: klt_level0_tma_kernel 157.97 MB read + 330.78 MB written, klt_levels_fused_kernel 39.44 + 61.84 MB, 256 pairs;
# less than the 2 649 600 algorithmic bytes per pair because part of the output is still in L2 when the kernels end)
KLT_NCU_TRAFFIC_PER_PAIR = (157.97e6 + 330.78e6 + 39.44e6 + 61.84e6) / 256


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the GPU is under the benchmark load."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill(); out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_setup(gpus: int):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int, device: str = "cuda") -> float:
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(x: float, world: int, device: str = "cuda") -> float:
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


# ------------------------------------------------------------------------------------------------
def bench_ekf(args, rank, world, local, comm=None):
    import torch
    from ekf_vio_b200 import capi, workload
    F, n, K, W = args.filters, N_FEAT, args.steps, args.warmup
    total_steps = K + W
    t0 = time.time()
    # Config 3 is 100 steps (the default: 97 timed + 3 warm-up).  Much longer runs of these streams drive a few
    # of the reference's filters unstable (their covariance grows without bound on any implementation); such filters end
    # with a non-zero status word and are reported in `checks`, not hidden.
    init_uv, meas, truth = workload.ekf_streams(rank * F, F, n, total_steps + 1, dt=DT)     # + one untimed step for the per-step oracle check
    gen_s = time.time() - t0
    h_meas = torch.from_numpy(meas).pin_memory()                  # [steps, F, n, 2]
    R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1))
    passed = np.ones((F, n), np.uint8)
    d_meas = h_meas.cuda(non_blocking=True)
    d_R = torch.from_numpy(R).cuda(); d_pass = torch.from_numpy(passed).cuda()
    d_truth = torch.from_numpy(truth[total_steps - 1]).cuda()
    d_acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    kvec = np.full(F, n, np.int32)

    batch = capi.EkfBatch(F, n, device=local)
    batch.add_features_h(kvec, init_uv)

    def run(steps_from, steps_to):
        for s in range(steps_from, steps_to):
            batch.process(DT)
            batch.update(d_meas[s], d_R, d_pass)

    # ---- device-resident arm ----
    sampler = ClockSampler(local)
    sampler.start()               # 50 ms samples across warm-up + timed region (the same load)
    run(0, W)
    barrier(world)
    batch.enable_timing(True)
    l0 = batch.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    run(W, W + K)
    batch.accumulate_errors(d_truth, d_acc)
    if comm is not None:
        comm.allreduce(d_acc[:4])                                 # Monte-Carlo error statistics: ekfvio_stats_allreduce = one ncclAllReduce over NVLink
    e1.record()
    barrier(world)
    ms_total = max_over_ranks(e0.elapsed_time(e1), world)
    launches = batch.launches - l0
    kms, kcnt = batch.timing()
    batch.enable_timing(False)
    acc = d_acc.cpu().numpy()
    st = batch.get_state(want_P=False)
    bad = int(((st["status"] & 3) != 0).sum())                          # zero pivot (the reference's NumericalIssue) or non-finite state
    ldlt = int(((st["status"] & 8) != 0).sum())                         # informational: S was not positive definite at some step, LDL^T kernels used
    finite = bool(np.isfinite(st["mu"][(st["status"] & 3) == 0]).all())
    # spot check of the timed batch itself against the FP64 oracle (the checker, not the thing measured): filters 0, F/3, 2F/3,
    # F-1 of this rank.  Per step — the oracle seeded with the batch's state after the timed steps must reproduce one more
    # (untimed) step of the batch within 1e-9 — and free-running over all W+K steps beside the oracle's own sensitivity.
    spot = sorted({0, F // 3, (2 * F) // 3, F - 1})
    oracle = None
    if not args.skip_cpu:
        before = {f: batch.get_state_range(f, 1) for f in spot}
        batch.process(DT); batch.update(d_meas[total_steps], d_R, d_pass)
        after = {f: batch.get_state_range(f, 1) for f in spot}
        oracle = oracle_spot_check(before, after, spot, init_uv, meas, R, passed, n, total_steps)
    # ---- end-to-end arm: host buffers in, host state out, every step ----
    batch.reset()
    batch.add_features_h(kvec, init_uv)
    h_mu = torch.zeros(F, 22, dtype=torch.float64).pin_memory().numpy()          # page-locked: the state is DMA'd straight into them
    h_feat = torch.zeros(F, n, 3, dtype=torch.float64).pin_memory().numpy()
    meas_np = h_meas.numpy()                                   # page-locked: the C ABI DMAs straight from it
    R_pin = torch.from_numpy(R).pin_memory().numpy(); passed_pin = torch.from_numpy(passed).pin_memory().numpy()

    def run_e2e(a, b):
        for s in range(a, b):
            batch.process(DT)
            batch.update_h(meas_np[s], R_pin, passed_pin)
            batch.read_mu_h(h_mu, h_feat)

    run_e2e(0, W)
    barrier(world)
    t0 = time.perf_counter()
    run_e2e(W, W + K)
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    clocks = sampler.stop()      # sampled across both timed regions (device-resident and end-to-end arm: the same kernels)
    h2d = meas_np[0].nbytes + R.nbytes + passed.nbytes
    d2h = h_mu.nbytes + h_feat.nbytes
    # the two arms must agree on the final state (same stream, same arithmetic)
    st2 = batch.get_state(want_P=False)
    arms_agree = bool(np.array_equal(st["mu"], st2["mu"], equal_nan=True))

    total_filters = F * world
    value = total_filters * K / (ms_total * 1e-3)
    res = {
        "value": value, "ms_per_step": ms_total / K, "launches": int(launches), "clocks": clocks,
        "e2e": {"value": total_filters * K / e2e_s, "unit": "filter-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        # update_fused: ekf_update_fused; gain_chol / gain_solve / cov_update: the tiled kernels behind it, which serve the filters it leaves
        # alone (S not positive definite or ill conditioned: a handful per step) — their time is the latency of those few filters
        "kernel_ms": {"process": kms[0] / max(kcnt[0], 1), "update_fused": kms[4] / max(kcnt[4], 1), "gain_chol": kms[1] / max(kcnt[1], 1),
                      "gain_solve": kms[3] / max(kcnt[3], 1), "cov_update": kms[2] / max(kcnt[2], 1)},
        "kernel_share": {k: float(v) for k, v in zip(("process", "gain_chol", "cov_update", "gain_solve", "update_fused"), kms[:5] / max(kms[:5].sum(), 1e-12))},
        "status_nonzero": bad, "ldlt_filters": ldlt, "finite": finite, "arms_agree": arms_agree, "gen_s": gen_s, "oracle": oracle, "oracle_filters": spot,
        "mc_stats": {"rmse_pos": float(np.sqrt(acc[0] / max(acc[3], 1))), "rmse_vel": float(np.sqrt(acc[1] / max(acc[3], 1))), "count": float(acc[3])},
    }
    batch.close()
    return res


def oracle_spot_check(before, after, filters, init_uv, meas, R, passed, n, steps):
    """bench.py may execute oracle/ only as the checker.  Max-norm relative errors (state and Sigma) of one step of the timed
    batch — step number `steps`, taken untimed after the timed region — on the given filters:
      one_step     batch vs the FP64 oracle (oracle/ekf_oracle.hpp) seeded with the batch's own state before the step (gate 1e-9,
                   north_star's "per step");
      oracle_own   the FP64 oracle vs the same step evaluated in 80-bit extended precision: the rounding error the reference's
                   step itself carries in FP64.  On the ill-conditioned steps of these streams (cond(S) up to 1e15) it is far
                   above 1e-9 and one_step cannot be below it;
      batch_true   batch vs the extended-precision result."""
    from tests import oracle_lib as O
    N = 22 + 3 * n
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    full = lambda s: np.concatenate([s["mu"], s["feat"].ravel()])
    one = own = true = 0.0
    per_filter = []
    for f in filters:
        b0, a0 = before[f], after[f]
        o = O.OracleFilter(); o.add_features(init_uv[f])
        o.set_state(mu=b0["mu"][0], feat=b0["feat"][0, :n], Pm=b0["P"][0, :N, :N], cache=b0["cache"][0], flags=b0["flags"][0, :n], klt_last=b0["klt_last"][0, :n])
        o.process(DT); o.update(meas[steps, f], R[f], passed[f])
        s1 = o.state()
        ex = O.step_extended(b0["mu"][0], b0["feat"][0, :n], b0["P"][0, :N, :N], b0["cache"][0], DT, meas[steps, f], R[f], passed[f])
        g1 = np.concatenate([a0["mu"][0], a0["feat"][0, :n].ravel()]); gP = a0["P"][0, :N, :N]
        e1 = max(rel(g1, full(s1)), rel(gP, s1["P"])); e2 = max(rel(full(s1), full(ex)), rel(s1["P"], ex["P"]))
        one = max(one, e1); own = max(own, e2)
        true = max(true, rel(g1, full(ex)), rel(gP, ex["P"]))
        # the gate of tests/test_gpu_ekf.py::test_config3_stream_100_steps: 1e-9 wherever FP64 carries the reference's step that far, else
        # 100 x the rounding error the FP64 oracle itself has on this step
        per_filter.append({"filter": int(f), "one_step": e1, "oracle_own": e2, "gate": max(1e-9, 100 * e2), "pass": bool(e1 <= max(1e-9, 100 * e2))})
    healthy = [x["one_step"] for x in per_filter if x["oracle_own"] <= 1e-10]
    return {"one_step": one, "oracle_own": own, "batch_true": true, "per_filter": per_filter,
            "one_step_well_conditioned": max(healthy) if healthy else None}


def flops_update_executed(n: int) -> float:
    """Reduced update as executed (symmetric counting): N m^2 + N^2 m + m^3/3."""
    N, m = 22 + 3 * n, 2 * n
    return 1.0 * N * m * m + 1.0 * N * N * m + m ** 3 / 3


def bench_ekf_rate(F, n, K, W, rank, world, local, comm, oracle_check):
    """Device-resident rate of one configuration (config 4: large states; config 5: the 65 536-filter sweep), same step as
    bench_ekf: process(dt) + update(all n measured), streams keyed on the global filter index.  Optionally one step of filter 0
    against the FP64 oracle, seeded with the batch's own state."""
    import torch
    from ekf_vio_b200 import capi, workload
    init_uv, meas, truth = workload.ekf_streams(rank * F, F, n, K + W + 1, dt=DT)
    R = np.tile(np.array([1e-5, 0, 0, 1e-5]), (F, n, 1)); passed = np.ones((F, n), np.uint8)
    d_meas = torch.from_numpy(meas).cuda(); d_R = torch.from_numpy(R).cuda(); d_pass = torch.from_numpy(passed).cuda()
    d_truth = torch.from_numpy(truth[K + W - 1]).cuda(); d_acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    batch = capi.EkfBatch(F, n, device=local)
    batch.add_features_h(np.full(F, n, np.int32), init_uv)
    for s in range(W):
        batch.process(DT); batch.update(d_meas[s], d_R, d_pass)
    barrier(world)
    l0 = batch.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        batch.process(DT); batch.update(d_meas[s], d_R, d_pass)
    batch.accumulate_errors(d_truth, d_acc)
    if comm is not None:
        comm.allreduce(d_acc[:4])
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = batch.launches - l0
    acc = d_acc.cpu().numpy()
    st = batch.get_state(want_P=False)
    res = {"filters_per_gpu": F, "filters_total": F * world, "features": n, "steps": K, "warmup": W, "value": F * world * K / (ms * 1e-3), "unit": "filter-steps/s",
           "ms_per_step": ms / K, "gpu_launches": int(launches), "status_nonzero": int(((st["status"] & 3) != 0).sum()), "finite": bool(np.isfinite(st["mu"]).all()),
           "mc_stats": {"rmse_pos": float(np.sqrt(acc[0] / max(acc[3], 1))), "rmse_vel": float(np.sqrt(acc[1] / max(acc[3], 1))), "count": float(acc[3])}}
    if oracle_check:
        from tests import oracle_lib as O
        N = 22 + 3 * n
        b0 = batch.get_state_range(0, 1)
        batch.process(DT); batch.update(d_meas[K + W], d_R, d_pass)
        a0 = batch.get_state_range(0, 1)
        o = O.OracleFilter(); o.add_features(init_uv[0])
        o.set_state(mu=b0["mu"][0], feat=b0["feat"][0, :n], Pm=b0["P"][0, :N, :N], cache=b0["cache"][0])
        o.process(DT); o.update(meas[K + W, 0], R[0], passed[0])
        s1 = o.state()
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        res["oracle_rel"] = max(rel(np.concatenate([a0["mu"][0], a0["feat"][0, :n].ravel()]), np.concatenate([s1["mu"], s1["feat"].ravel()])), rel(a0["P"][0, :N, :N], s1["P"]))
        res["oracle_tol"] = 1e-9
    batch.close()
    return res


def bench_replenish(args, rank, world, local, prev, pts):
    """SURVEY.md §8(f)1 — EKFVIO::replenishFeatures for a batch of frames: cv::FAST(50, nms) + check image + greedy
    scan with 60 features already in the state and 40 wanted (NUM_FEATURES 100).  Frames resident in HBM; e2e from host."""
    import torch
    from ekf_vio_b200 import capi
    B, K, W = prev.shape[0], args.steps, args.warmup
    det = capi.FastDetector(640, 480, B, 4096, device=local)
    d = torch.from_numpy(prev).cuda()
    kp = torch.zeros(B, 4096, 2, dtype=torch.int16, device="cuda"); cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    needed = torch.full((B,), 40, dtype=torch.int32, device="cuda")
    ex_h = np.ascontiguousarray(pts[:, :60]); nex_h = np.full(B, 60, np.int32)
    ex = torch.from_numpy(ex_h).cuda(); nex = torch.from_numpy(nex_h).cuda()
    new_px = torch.zeros(B, 128, 2, dtype=torch.int16, device="cuda"); n_new = torch.zeros(B, dtype=torch.int32, device="cuda")

    def step():
        det.detect(d, 50, True, kp, None, cnt)
        det.select(kp, cnt, ex, nex, needed, 30, 11, None, new_px, None, n_new)

    for _ in range(W):
        step()
    barrier(world)
    l0 = det.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(K):
        step()
    e1.record(); torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = det.launches - l0
    frames = sum_over_ranks(float(B), world)
    t0 = time.perf_counter()
    for _ in range(K):
        det.replenish_h(prev, 50, ex_h, nex_h, np.full(B, 40, np.int32))
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    res = {"metric": "frames replenished/s at 640x480 (FAST-9/16 + NMS + greedy min-distance scan)", "value": frames * K / (ms * 1e-3), "unit": "frames/s",
           "ms_per_step": ms / K, "keypoints_per_frame": float(cnt.float().mean().item()), "new_per_frame": float(n_new.float().mean().item()),
           "gpu_launches": int(launches), "e2e": {"value": frames * K / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(prev.nbytes + ex_h.nbytes)}}
    det.close()
    return res


def bench_vio_loop(args, rank, world, local):
    """SURVEY.md §8(f)3 — EKFVIO::addFrame on the device: one call advances S sequences by one 640x480 frame
    (pyramid, process, KLT on <= 100 features, update, FAST replenishment).  Frames resident in HBM."""
    import torch
    from ekf_vio_b200 import capi, workload
    S = 64
    warm = max(args.warmup + 1, 5)            # frames 0-4: eager frames and the two graph captures (one per pyramid-slot parity)
    T = warm + 6
    frames = workload.vio_sequences(rank * S, 8, T, 640, 480, speed=2.0)
    frames = np.ascontiguousarray(np.concatenate([frames] * (S // 8), axis=1))
    K9 = np.zeros((S, 9), np.float32); K9[:, 0] = 400.0; K9[:, 4] = 400.0; K9[:, 6] = 320.0; K9[:, 7] = 240.0; K9[:, 8] = 1.0
    loop = capi.VioLoop(S, 640, 480, device=local)
    d_frames = torch.from_numpy(frames).cuda(); dK = torch.from_numpy(K9).cuda(); ddt = torch.full((S,), DT, dtype=torch.float64, device="cuda")
    for t in range(warm):
        loop.add_frame(d_frames[t], dK, None if t == 0 else ddt)
    barrier(world)
    l0 = loop.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for t in range(warm, T):
        loop.add_frame(d_frames[t], dK, ddt)
    e1.record(); torch.cuda.synchronize()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    st = loop.filters.get_state(want_P=False)
    res = {"metric": "sequence-frames/s through the device frame loop (640x480, NUM_FEATURES 100)", "value": sum_over_ranks(float(S), world) * (T - warm) / (ms * 1e-3),
           "unit": "frames/s", "ms_per_frame_step": ms / (T - warm), "sequences_per_gpu": S, "gpu_launches": int(loop.launches - l0),
           "features_per_sequence": float(st["nfeat"].mean()), "status_nonzero": int((st["status"] != 0).sum())}
    loop.close()
    return res


def bench_klt(args, rank, world, local):
    """KLT features tracked/s at 640x480.  Headline leg: SURVEY.md §8d's synthetic rate workload — config 2's base image warped by
    seeded random affine maps (|t| <= 24 px, |shear| <= 0.05), points = its first 200 FAST(50, nms) corners.  Second leg: config 2
    itself (test -> moved / test -> shear), whose tracked counts must be cv2's (190 / 194)."""
    import torch
    from ekf_vio_b200 import capi, workload
    B, npts, K, W = args.klt_pairs, 200, args.steps, args.warmup
    G = min(B, 32)                      # distinct pairs per rank; the batch cycles through them (separate buffers: nothing is shared in L2)
    reps = (B + G - 1) // G
    tile = lambda a: np.ascontiguousarray(np.concatenate([a] * reps)[:B])
    prev, nxt, pts, flow = (tile(a) for a in workload.klt_pairs_8d(rank * G, G))
    trk = capi.KltTracker(640, 480, B, npts, device=local)
    d_out = torch.zeros(B, npts, 2, dtype=torch.float32, device="cuda"); d_status = torch.zeros(B, npts, dtype=torch.uint8, device="cuda")
    d_err = torch.zeros(B, npts, dtype=torch.float32, device="cuda")
    d_npts = torch.full((B,), npts, dtype=torch.int32, device="cuda")
    K9 = np.zeros((B, 9), np.float32); K9[:, 0] = 400.0; K9[:, 4] = 400.0; K9[:, 8] = 1.0
    d_K9 = torch.from_numpy(K9).cuda()
    d_meas = torch.zeros(B, npts, 2, dtype=torch.float32, device="cuda"); d_cov = torch.zeros(B, npts, 4, dtype=torch.float32, device="cuda")
    d_passed = torch.zeros(B, npts, dtype=torch.uint8, device="cuda")

    def device_leg(prev_h, next_h, pts_h, steps, warm):
        d_prev = torch.from_numpy(prev_h).cuda(); d_next = torch.from_numpy(next_h).cuda(); d_pts = torch.from_numpy(pts_h).cuda()

        def step():
            # both pyramids, as cv::calcOpticalFlowPyrLK rebuilds them per call; level 0 is the caller's image (no copy), as it is there
            trk.build_pyramid_pair_ref(0, d_prev, 1, d_next, False)
            d_out.copy_(d_pts)                                        # initial flow = previous positions (OPTFLOW_USE_INITIAL_FLOW)
            trk.track(0, 1, d_pts, d_out, d_status, d_err, d_npts)
            trk.postprocess(d_out, d_status, d_npts, d_K9, d_meas, d_cov, d_passed)

        for _ in range(warm):
            step()
        barrier(world)
        trk.enable_timing(True)
        l0 = trk.launches
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier(world)
        ms_total = max_over_ranks(e0.elapsed_time(e1), world)
        launches = trk.launches - l0
        kms, kcnt = trk.timing()
        trk.enable_timing(False)
        return ms_total, launches, kms, kcnt, d_pts

    ms_total, launches, kms, kcnt, d_pts = device_leg(prev, nxt, pts, K, W)
    tracked = float(d_passed.sum().item())
    attempted = float(B * npts)
    tracked_all = sum_over_ranks(tracked, world); attempted_all = sum_over_ranks(attempted, world)
    good = d_passed.bool()
    fl = (d_out - d_pts)[good].view(-1, 2)
    # sanity: the recovered flow of a pair is its affine map's (translation + shear * (y - 240) in x)
    exp = torch.from_numpy(flow).cuda()[:, None, :].expand(B, npts, 2)[good].view(-1, 2)
    flow_err_y = float((fl[:, 1] - exp[:, 1]).abs().median().item()) if fl.numel() else None

    # e2e: host images + points in (page-locked host memory), host results out
    prev_p = torch.from_numpy(prev).pin_memory().numpy(); nxt_p = torch.from_numpy(nxt).pin_memory().numpy()
    nn = pts.copy()
    npts_h = np.full(B, npts, np.int32)
    for _ in range(max(1, W // 2)):
        nn[:] = pts; trk.track_pair_h(prev_p, nxt_p, pts, nn, npts_h)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(K):
        nn[:] = pts
        st_h, _ = trk.track_pair_h(prev_p, nxt_p, pts, nn, npts_h)
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    in_pad = ~((nn[..., 0] < 11) | (nn[..., 1] < 11) | (640 - nn[..., 0] < 11) | (480 - nn[..., 1] < 11))
    tracked_e2e = sum_over_ranks(float(((st_h == 1) & in_pad).sum()), world)
    e2e = {"value": tracked_e2e * K / e2e_s, "unit": "features/s", "h2d_bytes_per_step": int(prev.nbytes + nxt.nbytes + 2 * pts.nbytes + npts_h.nbytes),
           "d2h_bytes_per_step": int(nn.nbytes + st_h.nbytes + 4 * st_h.size),
           "note": "ekfvio_klt_track_pair_h: both frames of every pair uploaded per call, as cv::calcOpticalFlowPyrLK takes them"}
    # a tracker in a sequence only needs the new frame: the previous frame's pyramid (with derivatives) is already on the device
    if hasattr(trk, "track_next_h"):
        trk.track_pair_h(prev_p, nxt_p, pts, nn, npts_h)                  # establishes "previous"
        seq = [nxt_p, prev_p]
        for i in range(2):
            nn[:] = pts; trk.track_next_h(seq[i % 2], pts, nn, npts_h)
        barrier(world)
        t0 = time.perf_counter()
        tr = 0.0
        for i in range(K):
            nn[:] = pts
            st_n, _ = trk.track_next_h(seq[i % 2], pts, nn, npts_h)
        barrier(world)
        seq_s = max_over_ranks(time.perf_counter() - t0, world)
        in_pad = ~((nn[..., 0] < 11) | (nn[..., 1] < 11) | (640 - nn[..., 0] < 11) | (480 - nn[..., 1] < 11))
        e2e["sequence"] = {"value": sum_over_ranks(float(((st_n == 1) & in_pad).sum()), world) * K / seq_s, "unit": "features/s",
                           "h2d_bytes_per_step": int(nxt.nbytes + 2 * pts.nbytes + npts_h.nbytes),
                           "note": "ekfvio_klt_track_next_h: only the new frame travels; the previous pyramid stays on the device (KLTTracker in EKFVIO::addFrame)"}

    # roofline of the pyramid + Scharr construction (all levels of both images of a pair)
    peaks = load_peaks()
    pyr_ms = float(kms[:4].sum() / K)
    pyr_bytes = B * (KLT_BYTES_WITH_DERIVS + KLT_BYTES_NO_DERIVS)
    bytes_l0 = B * ((640 * 480 + 640 * 480 * 4 + 320 * 240) + (640 * 480 + 320 * 240))
    ms_l0 = kms[0] / max(kcnt[0], 1)
    res = {
        "metric": "KLT features tracked/s at 640x480", "value": tracked_all * K / (ms_total * 1e-3), "unit": "features/s",
        "attempted_per_s": attempted_all * K / (ms_total * 1e-3), "ms_per_step": ms_total / K,
        "config": {"workload": f"{B} image pairs/GPU 640x480 (SURVEY 8d: config 2's base image under seeded affine maps, |t| <= 24 px, |shear| <= 0.05; {G} distinct pairs/GPU), "
                               "its first 200 FAST(50, nms) corners per pair, both pyramids rebuilt per step (as cv::calcOpticalFlowPyrLK does), win 21, 4 levels; inputs larger than L2"},
        "tracked_fraction": tracked / attempted, "tracked": tracked_all, "attempted": attempted_all, "median_flow_err_px": flow_err_y,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "kernel_ms": {"pyr_level0": ms_l0, "pyr_levels_1_3": float(kms[1:4].sum() / max(kcnt[1], 1)), "pyramids_per_step": pyr_ms, "track": kms[4] / max(kcnt[4], 1)},
        "roofline": {"bound": "hbm", "kernel": "klt_level0_tma_kernel + klt_levels_fused_kernel (pyramid + Scharr, all levels of both images; TMA-staged)",
                     "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "bytes_per_launch_pair": pyr_bytes, "traffic": KLT_NCU_TRAFFIC_PER_PAIR * B if KLT_NCU_TRAFFIC_PER_PAIR else None,
                     "peak_source": peaks["source"], "level0_achieved": bytes_l0 / (ms_l0 * 1e-3) / 1e9},
    }
    # ---- config 2 itself: test -> moved / test -> shear, 200 FAST corners (BASELINE.json configs[1]) ----
    p2, n2, pts2, expect = workload.klt_pairs_config2(B)
    ms2, _, _, _, _ = device_leg(p2, n2, pts2, max(3, K // 4), 2)
    st2 = d_status.cpu().numpy()
    per_pair = st2.sum(1)
    res["config2"] = {"workload": f"{B} pairs/GPU: images/640_480_test.png -> moved / shear variants (alternating), first 200 FAST(50, nms) corners",
                      "value": sum_over_ranks(float(d_passed.sum().item()), world) * max(3, K // 4) / (ms2 * 1e-3), "unit": "features/s",
                      "ms_per_step": ms2 / max(3, K // 4), "status_tracked_per_pair": [int(per_pair[0]), int(per_pair[1])], "cv2_tracked_per_pair": [int(expect[0]), int(expect[1])],
                      "status_counts_match_cv2": bool((per_pair == expect).all()), "passed_kill_pad_fraction": float(d_passed.float().mean().item())}
    trk.close()
    res["replenish"] = bench_replenish(args, rank, world, local, prev, pts)
    res["vio_loop"] = bench_vio_loop(args, rank, world, local)
    return res


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"hbm_gbs": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


# ------------------------------------------------------------------------------------------------
def cpu_baseline_ekf(sample_filters: int, sample_steps: int, threads: int = 0):
    """Times the FP64 oracle (a port; DESIGN.md section 3 says why oracle/_ref is not the timed arm)
    on the host cores, OpenMP over filters.  bench.py's cpu_baseline leg is one of the two places
    allowed to execute oracle/."""
    import ctypes as C
    from ekf_vio_b200 import workload
    so = os.path.join(ROOT, "oracle", "libekf_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.ekfo_batch_run.restype = C.c_double
    lib.ekfo_max_threads.restype = C.c_int
    cores = threads or (os.cpu_count() or 1)
    init_uv, meas, _ = workload.ekf_streams(0, sample_filters, N_FEAT, sample_steps, dt=DT)
    init_uv = np.ascontiguousarray(init_uv); meas = np.ascontiguousarray(meas)
    sec = lib.ekfo_batch_run(C.c_int(sample_filters), C.c_int(N_FEAT), C.c_int(sample_steps), C.c_double(DT), init_uv.ctypes.data_as(C.c_void_p), None,
                             meas.ctypes.data_as(C.c_void_p), C.c_double(1e-5), C.c_int(cores), None, None)
    return {"value": sample_filters * sample_steps / sec, "unit": "filter-steps/s", "cores": int(cores), "kind": "port",
            "sample": f"{sample_filters} filters x {sample_steps} steps of the same workload (n=50), oracle/ekf_oracle.hpp, OpenMP over filters, {sec:.2f} s wall"}


def cpu_baseline_klt(pairs: int = 8, reps: int = 3):
    """cv2.calcOpticalFlowPyrLK — the reference's actual arithmetic — on all host cores."""
    from ekf_vio_b200 import workload
    try:
        import cv2
    except ImportError:
        return {"value": None, "unit": "features/s", "cores": 0, "kind": "reference", "sample": "cv2 not importable on this box"}
    prev, nxt, pts, _ = workload.klt_pairs_8d(0, pairs)
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)

    def once():
        tr = 0
        for i in range(pairs):
            p0 = pts[i].reshape(-1, 1, 2)
            nx, st, _ = cv2.calcOpticalFlowPyrLK(prev[i], nxt[i], p0, p0.copy(), winSize=(21, 21), maxLevel=3, criteria=crit,
                                                 flags=cv2.OPTFLOW_USE_INITIAL_FLOW, minEigThreshold=1e-4)
            x, y = nx[:, 0, 0], nx[:, 0, 1]
            tr += int(((st[:, 0] == 1) & ~((x < 11) | (y < 11) | (640 - x < 11) | (480 - y < 11))).sum())
        return tr
    once()
    t0 = time.perf_counter()
    tracked = 0
    for _ in range(reps):
        tracked += once()
    sec = time.perf_counter() - t0
    # replenishFeatures' detector on the same frames (cv::FAST(50, nms); the scalar scan behind it is negligible)
    fd = cv2.FastFeatureDetector_create(threshold=50, nonmaxSuppression=True)
    fd.detect(prev[0])
    t1 = time.perf_counter()
    for _ in range(reps):
        for i in range(pairs):
            fd.detect(prev[i])
    fast_fps = reps * pairs / (time.perf_counter() - t1)
    return {"value": tracked / sec, "unit": "features/s", "cores": int(cores), "kind": "reference",
            "sample": f"cv2 {cv2.__version__} calcOpticalFlowPyrLK, {pairs} pairs x {reps} reps, 200 points each, {cores} threads",
            "replenish_frames_per_s": fast_fps}


def run_reference(args, result_out=sys.stdout):
    """--impl reference: the reference's CPU path for the same metric/config, on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    cores = os.cpu_count() or 1
    # each "step" is a bounded sample of the 4096-filter batch: sample_filters filters stepped once
    sample_filters = max(cores * 4, 64)
    cb = cpu_baseline_ekf(sample_filters, max(1, min(W, 2)))          # warm-up
    t = cpu_baseline_ekf(sample_filters, K)
    v = t["value"]
    kl = cpu_baseline_klt(pairs=8, reps=max(1, min(K, 5)))
    line = {
        "impl": "reference", "metric": "EKF filter-steps/s", "value": v, "unit": "filter-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * args.filters * max(args.gpus, 1) / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {**ekf_config(args.filters, N_FEAT, max(args.gpus, 1)),
                   "note": "FP64 oracle port timed on a bounded sample (OpenMP over filters); oracle/_ref, the reference's own sources built against stand-in Eigen headers, is the parity pin, not the timed arm: its process-wide static cache (E2) rules out threads and its speed would be the shim's, not Eigen's"},
        "cpu_baseline": {**t, "value": v},
        "e2e": {"value": v, "unit": "filter-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "klt": {"metric": "KLT features tracked/s at 640x480", "value": kl["value"], "unit": "features/s", "cpu_baseline": kl},
    }
    del cb
    print(json.dumps(line), file=result_out, flush=True)


def claim_stdout():
    """stdout must carry exactly one JSON line, but libraries write to fd 1 from C (NCCL's version banner at
    communicator creation): everything is sent to stderr and the result line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    result_out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=97)   # + 3 warm-up = the 100 steps of config 3
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filters", type=int, default=4096, help="EKF filters per GPU")
    ap.add_argument("--klt-pairs", type=int, default=256, help="image pairs per GPU for the KLT leg")
    ap.add_argument("--skip-klt", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true", help="skip the config-4 (n=300) and config-5 (65 536 filters, N>1) legs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, result_out)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = dist_setup(args.gpus)
    from ekf_vio_b200 import capi

    dmma_peak, dfma_peak = capi.measure_fp64_peak(local)
    comm = None
    if world > 1:
        # the library's own collective (ekfvio_stats_allreduce, NCCL): torch.distributed only carries the 128-byte id to the ranks
        import torch.distributed as dist

        def exchange(raw):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())
        comm = capi.StatsComm(local, world, rank, exchange)
        warm = torch.zeros(4, dtype=torch.float64, device="cuda")
        comm.allreduce(warm); torch.cuda.synchronize()             # NCCL connects lazily on the first collective: keep that out of the timed regions
    ekf = bench_ekf(args, rank, world, local, comm)
    klt = None if args.skip_klt else bench_klt(args, rank, world, local)
    # config 4 (BASELINE.json configs[3]): 64 filters/GPU x (22 + 3*300) states, the blocked multi-CTA DMMA path
    cfg4 = None if args.skip_large else bench_ekf_rate(64, 300, min(args.steps, 12), 3, rank, world, local, comm, oracle_check=(rank == 0 and not args.skip_cpu))
    # config 5 (BASELINE.json configs[4]): 65 536 filters in total, sharded over the ranks (only when there are ranks to shard over)
    cfg5 = bench_ekf_rate(65536 // world, N_FEAT, min(args.steps, 10), 3, rank, world, local, comm, oracle_check=False) if world > 1 and not args.skip_large else None

    if rank == 0:
        n = N_FEAT
        fstep = flops_per_filter_step(n)
        F = args.filters
        peak = max(dmma_peak, dfma_peak)
        fused = ekf["kernel_ms"]["update_fused"] > 0
        cov_ms = ekf["kernel_ms"]["update_fused"] if fused else ekf["kernel_ms"]["cov_update"]
        cov_flops = flops_update_fused_executed(n) if fused else flops_cov_update_executed(n)
        cov_ref_flops = (flops_per_filter_step(n) - (486 * n * n + 5.7e3 * n + 4.3e4 + (1.75e3 * n + 5.3e3))) if fused else flops_cov_update(n)
        cov_achieved = F * cov_flops / (cov_ms * 1e-3) / 1e12 if cov_ms > 0 else 0.0
        cov_reference = F * cov_ref_flops / (cov_ms * 1e-3) / 1e12 if cov_ms > 0 else 0.0
        step_achieved = (ekf["value"] / world) * fstep / 1e12
        step_executed = (ekf["value"] / world) * flops_per_filter_step_executed(n) / 1e12
        line = {
            "metric": "EKF filter-steps/s", "value": ekf["value"], "unit": "filter-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ekf["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": ekf_config(F, n, world),
            "clocks": ekf["clocks"],
            "e2e": ekf["e2e"],
            "gpu_launches": ekf["launches"] + (klt["gpu_launches"] if klt else 0),
            # The dominant kernel is the whole measurement update, ekf_update_fused: `achieved` / `frac` count the FLOPs its
            # DMMAs really execute (flops_update_fused_executed — what the FP64 pipe sees); `reference_equivalent` rates the same launch in
            # the FLOPs of the update as the reference writes it (4 N^2 m + 2 N m^2 + m^3/3).
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA.8x8x4 / DFMA share one pipe on sm_100)",
                         "kernel": "measurement update (ekf_update_fused)" if fused else "covariance update (ekf_joseph_sym)",
                         "achieved": cov_achieved, "peak": peak, "unit": "TFLOP/s", "frac": cov_achieved / peak if peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the kernel from this round's ncu --set full capture
                         # (profiles/r02f_ncu_ekf_summary.txt, 4096 filters), scaled to this launch's filter count;
                         # a profiler figure cannot be re-measured inside a timed run
                         "traffic": ((NCU_FUSED_DRAM_BYTES if fused else 1.2013e9 + 0.9253e9) / 4096 * F) if (n == 50 and (NCU_FUSED_DRAM_BYTES or not fused)) else None,
                         "peak_source": "measured live by ekfvio_measure_fp64_peak (register-resident DMMA/DFMA loops); MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_launch": F * cov_flops,
                         "reference_equivalent": {"flops_per_launch": F * cov_ref_flops, "achieved": cov_reference},
                         "whole_step": {"flops_per_filter_step": fstep, "achieved": step_achieved, "frac": step_achieved / peak if peak else None,
                                        "note": "SURVEY.md §8d algorithmic count of the reference's update (Joseph form)",
                                        "executed_flops_per_filter_step": flops_per_filter_step_executed(n), "executed_achieved": step_executed,
                                        "executed_frac": step_executed / peak if peak else None}},
            "kernel_ms": ekf["kernel_ms"], "kernel_share": ekf["kernel_share"],
            "fp64_peak_tflops": {"dmma": dmma_peak, "dfma": dfma_peak},
            # oracle_rel: one step of the timed batch against the FP64 oracle seeded with the batch's own state (gate 1e-9, north_star's
            # "per step"); oracle_own_error: the FP64 oracle against the same step in 80-bit extended precision — the floor FP64 itself
            # sets on ill-conditioned steps of these streams; batch_true_error: the batch against extended precision
            "checks": {"status_nonzero": ekf["status_nonzero"], "ldlt_filters": ekf["ldlt_filters"], "finite": ekf["finite"], "arms_agree": ekf["arms_agree"],
                       "oracle_rel": ekf["oracle"]["one_step"] if ekf["oracle"] else None, "oracle_tol": 1e-9,
                       "oracle_own_error": ekf["oracle"]["oracle_own"] if ekf["oracle"] else None,
                       "batch_true_error": ekf["oracle"]["batch_true"] if ekf["oracle"] else None,
                       # per spot filter, with the tests' gate max(1e-9, 100 x the oracle's own error on that step); oracle_rel_well_conditioned:
                       # the filters whose step FP64 itself carries to 1e-10 (gate 1e-9)
                       "oracle_per_filter": ekf["oracle"]["per_filter"] if ekf["oracle"] else None,
                       "oracle_rel_well_conditioned": ekf["oracle"]["one_step_well_conditioned"] if ekf["oracle"] else None,
                       "oracle_pass": all(x["pass"] for x in ekf["oracle"]["per_filter"]) if ekf["oracle"] else None,
                       "oracle_filters": ekf["oracle_filters"], "mc_stats": ekf["mc_stats"]},
        }
        if comm is not None:
            line["collective"] = {"api": "ekfvio_stats_allreduce (ncclAllReduce, FP64 sum, 4 doubles per report)", "nranks": comm.size()}
        if cfg4:
            f4 = flops_per_filter_step(300)
            fx4 = 486 * 300 * 300 + 5.7e3 * 300 + 4.3e4 + (1.75e3 * 300 + 5.3e3) + flops_update_executed(300)
            v4 = cfg4["value"] / world
            cfg4["workload"] = "BASELINE.json configs[3]: 64 filters/GPU x (22+3*300)-dim state (N = 922, m = 600), process + update per step, blocked multi-CTA DMMA path"
            cfg4["roofline"] = {"bound": "tensor", "pipe": "fp64 (DMMA.8x8x4)", "peak": peak, "unit": "TFLOP/s",
                                "survey_flops_per_filter_step": f4, "achieved": v4 * f4 / 1e12, "frac": v4 * f4 / 1e12 / peak if peak else None,
                                "executed_flops_per_filter_step": fx4, "executed_achieved": v4 * fx4 / 1e12, "executed_frac": v4 * fx4 / 1e12 / peak if peak else None}
            line["config4"] = cfg4
        if cfg5:
            v5 = cfg5["value"] / world
            cfg5["workload"] = "BASELINE.json configs[4]: 65 536 filters in total (n = 50) sharded over the ranks, error statistics reduced by ekfvio_stats_allreduce"
            cfg5["whole_step_frac_survey_flops"] = v5 * fstep / 1e12 / peak if peak else None
            line["config5"] = cfg5
        if not args.skip_cpu:
            line["cpu_baseline"] = cpu_baseline_ekf(256, 6)
        if klt:
            if not args.skip_cpu:
                klt["cpu_baseline"] = cpu_baseline_klt()
            line["klt"] = klt
        print(json.dumps(line), file=result_out, flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
