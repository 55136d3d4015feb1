"""GPU: bnn_moments_allreduce -- the NCCL collective inside the C ABI -- against the single-process moments and against the
all-gather + Chan-merge path.  Spawned as one process per GPU: world size 1 always (communicator plumbing + kernels), world size 2
when the box has two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from bnn_b200.distributed import NcclMoments, shard_range, merge_shards
    from bnn_b200 import ops
    nm = NcclMoments(dev)
    g = torch.Generator().manual_seed(0)
    S, M = 101, 1003                                             # M not a multiple of the world size
    pred = (5 + 0.01 * torch.randn(S, M, generator=g)).to(dev)   # the SURVEY 0.2 stress case, same on every rank
    lo, hi = shard_range(S, rank, world)
    _, _, pm, pm2 = ops.moments(pred[lo:hi].contiguous(), want_partials=True)
    mean, var = nm.merge(pm, pm2, hi - lo)
    mean2, var2 = merge_shards(pm, pm2, hi - lo)                 # all-gather + Chan merge kernel
    ref_mean, ref_var = pred.double().mean(0), pred.double().var(0)
    sl_mean, sl_var = nm.merge(pm, pm2, hi - lo, gather=False)
    Ms = (M + world - 1) // world
    a, b = rank * Ms, min(M, (rank + 1) * Ms)
    torch.cuda.synchronize()
    q.put((rank,
           float((mean.double() - ref_mean).abs().max() / ref_mean.abs().max()),
           float((var.double() - ref_var).abs().max() / ref_var.abs().max()),
           float((mean - mean2).abs().max()), float((var - var2).abs().max() / var2.abs().max()),
           float((sl_var[a:b].double() - ref_var[a:b]).abs().max() / ref_var.abs().max())))
    dist.barrier()
    nm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_moments_allreduce_matches_single_process(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, e_mean, e_var, d_mean, d_var, e_slice in res:
        assert e_mean <= 1e-6 and e_var <= 1e-5, (rank, e_mean, e_var)           # float32 outputs of float64 arithmetic
        assert d_mean <= 1e-6 and d_var <= 1e-5, (rank, d_mean, d_var)           # == the Chan-merge path
        assert e_slice <= 1e-5
