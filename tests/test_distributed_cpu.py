"""CPU, gloo, world_size 2: the N > 1 host logic -- sample sharding + all-gather of the per-point
float64 moments + Chan merge -- reproduces the single-process moments.  The merge kernel is CUDA-only,
so the test injects the oracle's merge for the arithmetic (bnn_b200.distributed.merge_shards(merge_fn=...))."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bnn_oracle as O


def _oracle_merge(pm, pm2, counts):
    parts = [(int(c), pm[r].numpy(), pm2[r].numpy()) for r, c in enumerate(counts)]
    n, mean, m2 = O.merge_moments(parts)
    return torch.from_numpy(mean).float(), torch.from_numpy(m2 / (n - 1)).float()


def _worker(rank, world, port, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bnn_b200.distributed import shard_range, merge_shards
    g = torch.Generator().manual_seed(0)
    pred = 5 + 0.01 * torch.randn(S, 37, generator=g)           # every rank builds the same full matrix
    lo, hi = shard_range(S, rank, world)
    n, mean, m2 = O.shard_moments(pred[lo:hi].double().numpy())
    mean_t, var_t = merge_shards(torch.from_numpy(mean), torch.from_numpy(m2), n, merge_fn=_oracle_merge)
    q.put((rank, lo, hi, mean_t.numpy(), var_t.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    S, world = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    [p.join(timeout=60) for p in procs]
    assert [(r[1], r[2]) for r in res] == [(0, 51), (51, 101)]
    g = torch.Generator().manual_seed(0)
    pred = 5 + 0.01 * torch.randn(S, 37, generator=g)
    for _, _, _, mean, var in res:
        assert np.allclose(mean, pred.mean(0).numpy(), rtol=1e-6)
        assert np.allclose(var, pred.var(0).numpy(), rtol=1e-5)
    assert np.array_equal(res[0][3], res[1][3]) and np.array_equal(res[0][4], res[1][4])   # identical on all ranks


def test_shard_range_covers_everything():
    from bnn_b200.distributed import shard_range
    for S in (1, 7, 8, 1250, 10000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(S, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
