"""World-size-2 test of the multi-process plumbing on CPU (gloo): shard assignment by global
index, max-over-ranks timing and the sum reduction of the Monte-Carlo statistics vector."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench
from ekf_vio_b200 import workload


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F = 3
    uv, meas, truth = workload.ekf_streams(rank * F, F, 4, 2)          # this rank's shard, as bench_ekf takes it
    acc = torch.tensor([float(np.square(uv).sum()), float(F), 0.0, 0.0], dtype=torch.float64)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    tmax = bench.max_over_ranks(10.0 + rank, world, device="cpu")
    tsum = bench.sum_over_ranks(float(rank + 1), world, device="cpu")
    # the KLT and frame-loop legs shard the same way: sequences by global index
    prev, nxt, pts, _ = workload.klt_pairs(rank * 2, 2, 128, 96, 5)
    seq = workload.vio_sequences(rank * 2, 2, 2, 96, 64)
    frames = torch.tensor([float(prev.shape[0])], dtype=torch.float64)
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    out[rank] = (acc.tolist(), tmax, tsum, uv.tolist(), int(prev.astype(np.int64).sum()), int(seq.astype(np.int64).sum()), float(frames[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    full_uv, _, _ = workload.ekf_streams(0, 6, 4, 2)
    np.testing.assert_array_equal(np.array(res[0][3] + res[1][3]), full_uv)      # shards tile the global index space
    full_prev = workload.klt_pairs(0, 4, 128, 96, 5)[0]
    full_seq = workload.vio_sequences(0, 4, 2, 96, 64)
    assert res[0][4] + res[1][4] == int(full_prev.astype(np.int64).sum())
    assert res[0][5] + res[1][5] == int(full_seq.astype(np.int64).sum())
    for r in range(world):
        acc, tmax, tsum = res[r][:3]
        assert abs(acc[0] - np.square(full_uv).sum()) < 1e-12 and acc[1] == 6.0
        assert tmax == 11.0 and tsum == 3.0 and res[r][6] == 4.0
