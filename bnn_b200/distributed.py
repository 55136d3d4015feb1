"""Multi-GPU plumbing: one process per GPU, posterior samples sharded across ranks.

The predictive path has exactly one exchange step: each rank reduces its own
samples to per-point float64 (mean, M2); the shards' moments are all-gathered over NCCL/NVLink
(16 bytes per point per rank) and Chan-merged on every rank by the bnn_moments_merge kernel, so
all ranks finish with identical mean / unbiased variance.  Raw float32 (sum, sum^2) is NOT used:
it loses ~13% of the variance when sigma << |mean|.
"""
import torch
import torch.distributed as dist


def shard_range(S, rank, world):
    """Contiguous, balanced [lo, hi) of S samples for `rank` (first S % world ranks get one more)."""
    base, extra = divmod(int(S), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_shards(part_mean, part_m2, count, group=None, merge_fn=None):
    """All-gather (count, mean, M2) of every rank and merge.  Returns (mean, var) float32.

    merge_fn(part_mean[R, ...], part_m2[R, ...], counts) -> (mean, var); defaults to the CUDA
    kernel (bnn_b200.ops.moments_merge).  There is no CPU implementation in this package; the
    gloo/CPU unit test injects the oracle's merge to exercise the sharding + collective logic.
    """
    if merge_fn is None:
        from . import ops
        merge_fn = ops.moments_merge
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_fn(part_mean.unsqueeze(0), part_m2.unsqueeze(0), [int(count)])
    R = dist.get_world_size(group)
    shape = tuple(part_mean.shape)
    packed = torch.cat([part_mean.reshape(-1), part_m2.reshape(-1)]).contiguous()      # [2*M] float64
    gathered = torch.empty(R * packed.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gathered = gathered.view(R, 2, -1)
    cnt = torch.tensor([int(count)], dtype=torch.int64, device=packed.device)
    cnts = torch.empty(R, dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    pm = gathered[:, 0].contiguous().view((R,) + shape)
    pm2 = gathered[:, 1].contiguous().view((R,) + shape)
    return merge_fn(pm, pm2, cnts.tolist())


# ---- the collective inside the C ABI (bnn_moments_allreduce) ---------------------------------------------------------------
import ctypes as _C
import glob as _glob
import os as _os


def _nccl_path():
    """The libnccl this process uses: the one bundled with torch (nvidia/nccl/lib), else the system's."""
    import sys
    for base in sys.path:
        hits = _glob.glob(_os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if hits:
            return hits[0]
    return "libnccl.so.2"


class _UniqueId(_C.Structure):
    _fields_ = [("internal", _C.c_ubyte * 128)]       # NOT c_char: ctypes truncates c_char arrays at the first NUL on read


class NcclMoments:
    """An NCCL communicator owned by this package (created with ncclCommInitRank from the process group's ranks; the unique id
    travels through torch.distributed) + the C-ABI collective that merges sample-sharded predictive moments:
    reduce-scatter of float64 raw moments, finalisation of this rank's slice of points, optional all-gather (see
    include/bnn_b200.h).  One instance per process; all calls are stream-ordered on torch's current stream."""

    def __init__(self, device, group=None):
        from . import _lib
        self.device = torch.device(device)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        path = _nccl_path()
        self._nccl = _C.CDLL(path, mode=_C.RTLD_GLOBAL)
        _lib.check(_lib.lib().bnn_nccl_load(path.encode()), "bnn_nccl_load")
        uid = _UniqueId()
        if self.rank == 0:
            rc = self._nccl.ncclGetUniqueId(_C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
        t = torch.frombuffer(bytearray(_C.string_at(_C.byref(uid), 128)), dtype=torch.uint8).clone()     # all 128 bytes
        t = t.to(self.device) if dist.get_backend(group) == "nccl" else t
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().numpy().tobytes())
        assert len(raw) == 128
        _C.memmove(_C.byref(uid), raw, 128)
        self.comm = _C.c_void_p()
        torch.cuda.set_device(self.device)
        self._nccl.ncclCommInitRank.argtypes = [_C.POINTER(_C.c_void_p), _C.c_int, _UniqueId, _C.c_int]
        rc = self._nccl.ncclCommInitRank(_C.byref(self.comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")
        self._ws = None

    def merge(self, part_mean, part_m2, count, gather=True):
        """(mean, var) float32 [..] from this rank's float64 shard moments; gather=False leaves only this rank's slice valid."""
        from . import _lib, ops
        if part_mean.dtype != torch.float64 or not part_mean.is_cuda:
            raise TypeError("shard moments must be float64 CUDA tensors")
        M = part_mean.numel()
        L = _lib.lib()
        need = int(L.bnn_moments_allreduce_workspace_bytes(M, self.world))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        mean = torch.empty(part_mean.shape, dtype=torch.float32, device=self.device)
        var = torch.empty(part_mean.shape, dtype=torch.float32, device=self.device)
        _lib.check(L.bnn_moments_allreduce(self.comm, ops._ptr(part_mean.contiguous()), ops._ptr(part_m2.contiguous()), int(count), M,
                                           ops._ptr(mean), ops._ptr(var), 1 if gather else 0, ops._ptr(self._ws), self._ws.numel(),
                                           ops._stream(mean)), "bnn_moments_allreduce")
        return mean, var

    def close(self):
        if self.comm:
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = _C.c_void_p()
