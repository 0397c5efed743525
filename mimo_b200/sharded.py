"""Data-sharded multi-GPU driver: one process per GPU, points split into contiguous shards,
ONE all-reduce per sweep of the packed FP64 statistics (+ the lower-bound scalar) over
NCCL / NVLink; every rank then runs the same deterministic batched posterior kernels, so no
broadcast is needed.  This is exactly the reference's list-of-arrays semantics -- per-shard
statistics are summed (distributions/gaussian.py:503-505, utils/abstraction.py:12-14).

Labels / uniforms / responsibilities stay sharded; Philox label draws are keyed by the
GLOBAL point index (Communicator.point_offset), so results do not depend on the shard count.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(N, rank, world):
    """contiguous [lo, hi) of `rank`; shard sizes differ by at most one point."""
    base, rem = divmod(N, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class Communicator:

    def __init__(self, N_global=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.N_global = N_global
        self.point_offset = shard_bounds(N_global, self.rank, self.world)[0] if N_global is not None else 0
        self.messages = 0
        self.bytes = 0

    def allreduce(self, t):
        """sum over shards, in place (FP64)."""
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self.messages += 1
        self.bytes += t.numel() * t.element_size()
        return t


class AbiCommunicator(Communicator):
    """Same interface, the all-reduce through the library's own C-ABI (mimo_comm_*: NCCL loaded by the library at run
    time) instead of torch.distributed -- what a caller that binds libmimo_b200.so without torch uses.  Here the 128-byte
    unique id of rank 0 is handed out over the already initialised torch.distributed group."""

    def __init__(self, N_global=None, group=None):
        super().__init__(N_global, group)
        import ctypes
        from . import _lib
        lib = _lib.load()
        nbytes = lib.mimo_comm_unique_id_bytes()
        buf = (ctypes.c_char * nbytes)()
        if self.rank == 0:
            _lib.call('mimo_comm_unique_id', ctypes.addressof(buf))
        box = [bytes(buf)]
        if self.world > 1:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._id = ctypes.create_string_buffer(box[0], nbytes)
        handle = ctypes.c_void_p()
        _lib.call('mimo_comm_init', self.world, self.rank, ctypes.addressof(self._id), ctypes.addressof(handle))
        self._handle = handle

    def allreduce(self, t):
        from . import _lib
        assert t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()
        _lib.call('mimo_comm_allreduce_stats', self._handle, t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream)
        self.messages += 1
        self.bytes += t.numel() * t.element_size()
        return t

    def close(self):
        from . import _lib
        if self._handle:
            _lib.call('mimo_comm_destroy', self._handle)
            self._handle = None


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    return int(os.environ.get('RANK', '0')), world
