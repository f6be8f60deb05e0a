"""Multi-GPU host logic: env partitioning and the one collective on the path.

The step shards by env with no cross-GPU traffic; the only exchange is the
RunningNorm statistics all-reduce (PHC/policies/running_norm.py:23-34 computed over the
concatenated batch).  Pure torch.distributed plumbing: works with NCCL on GPUs and gloo on CPU.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _cabi


def env_partition(num_envs: int, world_size: int, rank: int, multiple: int = 8) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) owned by ``rank``.  Boundaries fall on multiples of
    ``multiple`` (one kernel block) while they can, the last rank takes the remainder."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    blocks = -(-num_envs // multiple)
    per, extra = divmod(blocks, world_size)
    lo_b = rank * per + min(rank, extra)
    hi_b = lo_b + per + (1 if rank < extra else 0)
    return min(lo_b * multiple, num_envs), min(hi_b * multiple, num_envs)


def reduce_moments(sums: torch.Tensor, rows: int, group=None) -> torch.Tensor:
    """[sum(C) | sum of squares(C)] fp64 partials + local row count -> one all-reduce (SUM).
    Returns the payload ``[2C + 1]`` (last entry = total rows); identical on every rank."""
    if sums.dtype != torch.float64:
        raise TypeError("moment partials must be float64")
    payload = torch.cat([sums, torch.tensor([float(rows)], dtype=torch.float64, device=sums.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=group)
    return payload


class PeerReduce:
    """This rank's end of the fused all-reduce + RunningNorm blend over NVLink peer memory
    (``phc_running_norm_update_peers``, csrc/phc_peer.cu): one launch per rank replaces
    ``all_reduce`` + ``phc_running_norm_update``.  Ranks are processes (``connect`` with the all-gathered
    CUDA IPC handles) or, for tests and one-process-many-GPUs drivers, objects of one process
    (``connect_local``)."""

    def __init__(self, rank: int, world: int, cols: int, device, timeout_ms: int = 0):
        self.rank, self.world, self.cols = rank, world, cols
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _cabi.PhcError("PeerReduce needs a CUDA device (the gloo/CPU path is reduce_moments)")
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.load().phc_peer_reduce_create(rank, world, cols, timeout_ms, C.byref(self._handle)),
                        "phc_peer_reduce_create")  # fmt: skip

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(_cabi.PEER_HANDLE_BYTES)
        _cabi.check(_cabi.load().phc_peer_reduce_handle(self._handle, buf), "phc_peer_reduce_handle")
        return buf.raw

    def connect(self, handles: Sequence[bytes]) -> None:
        if len(handles) != self.world or any(len(h) != _cabi.PEER_HANDLE_BYTES for h in handles):
            raise _cabi.PhcError("connect() needs one 64-byte IPC handle per rank, in rank order")
        blob = C.create_string_buffer(b"".join(handles), self.world * _cabi.PEER_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.load().phc_peer_reduce_connect(self._handle, blob), "phc_peer_reduce_connect")

    @staticmethod
    def connect_local(ranks: List["PeerReduce"]) -> None:
        arr = (C.c_void_p * len(ranks))(*[r._handle for r in ranks])
        for r in ranks:
            _cabi.check(_cabi.load().phc_peer_reduce_connect_local(r._handle, arr), "phc_peer_reduce_connect_local")

    @classmethod
    def from_process_group(cls, cols: int, device, group=None, timeout_ms: int = 0) -> "PeerReduce":
        """One rank per process: exchange the IPC handles through ``torch.distributed`` (host plumbing only)."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        me = cls(rank, world, cols, device, timeout_ms)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, me.ipc_handle(), group=group)
        me.connect(handles)
        return me

    def update(self, running_mean, running_var, count, sums: torch.Tensor, rows) -> None:
        """Blend this rollout's statistics of ALL ranks into the running buffers; ``sums`` (this rank's fp64
        ``[2*cols]`` partials) is zeroed for the next rollout.  ``rows``: number or fp64 device scalar."""
        if sums.dtype != torch.float64 or sums.numel() != 2 * self.cols or not sums.is_cuda:
            raise _cabi.PhcError("sums must be a CUDA float64 tensor of 2*cols elements")
        rows_dev = rows.data_ptr() if isinstance(rows, torch.Tensor) else None
        _cabi.check(
            _cabi.load().phc_running_norm_update_peers(
                self._handle, sums.data_ptr(), rows_dev, 0.0 if rows_dev else float(rows), running_mean.data_ptr(),
                running_var.data_ptr(), count.data_ptr(), _cabi.stream_ptr(self.device),
            ),
            "phc_running_norm_update_peers",
        )  # fmt: skip

    def status(self) -> int:
        """Synchronises; number of completed reductions.  Raises if a launch timed out waiting for a peer."""
        done = C.c_int64()
        _cabi.check(_cabi.load().phc_peer_reduce_status(self._handle, C.byref(done)), "phc_peer_reduce_status")
        return done.value

    def resync(self, group=None, local: Optional[List["PeerReduce"]] = None) -> None:
        """Collective recovery after a timeout (``phc_peer_reduce_resync``): host barrier, every rank clears its
        mailbox / flags / epoch / status, host barrier.  ``local``: the ranks of one process (tests) instead of a
        process group.  Partials that a timed-out launch did not consume stay with the caller."""
        if local is not None:
            for r in local:
                with torch.cuda.device(r.device):
                    _cabi.check(_cabi.load().phc_peer_reduce_resync(r._handle), "phc_peer_reduce_resync")
            return
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if distributed:
            dist.barrier(group=group)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.load().phc_peer_reduce_resync(self._handle), "phc_peer_reduce_resync")
        if distributed:
            dist.barrier(group=group)

    def close(self):
        if getattr(self, "_handle", None):
            _cabi.load().phc_peer_reduce_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
