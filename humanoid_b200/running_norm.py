"""Observation normaliser with multi-GPU statistics (PHC/policies/running_norm.py).

``update(x)`` reproduces ``RunningNorm.update`` (:23-34): batch mean, biased variance,
blend with weight 1/count, count += 1.  The batch statistics come from per-column fp64
sum / sum-of-squares partials produced on the device (``phc_obs_moments`` over a rollout
buffer, or accumulated by the fused step's epilogue); when ``torch.distributed`` is
initialised the partials and the row count are summed with ONE all-reduce (NCCL over
NVLink on GPUs) — the only collective anywhere on the path — so every rank ends with the
statistics of the concatenated batch.  Buffers keep the reference's names, shapes and
dtypes so a ``state_dict`` stays interchangeable.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _cabi
from .parallel import reduce_moments


class RunningNorm(nn.Module):
    def __init__(self, shape: int, epsilon: float = 1e-5, clip: float = 10.0, device="cuda"):
        super().__init__()
        self.register_buffer("running_mean", torch.zeros((1, shape), dtype=torch.float32, device=device))
        self.register_buffer("running_var", torch.ones((1, shape), dtype=torch.float32, device=device))
        self.register_buffer("count", torch.ones(1, dtype=torch.float32, device=device))
        self.epsilon = epsilon
        self.clip = clip
        self.shape = shape
        self._peers = None  # parallel.PeerReduce once enable_peer_reduce() was called

    def __getstate__(self):
        """``torch.save(model)`` (running_norm.py:36-53): buffers and scalars travel, the peer mailbox (device memory
        mapped into other processes) does not — call ``enable_peer_reduce`` again after loading."""
        state = self.__dict__.copy()
        state["_peers"] = None
        state.pop("_payload", None)
        return state

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # :15-20
        _cabi.require_cuda(x, "x", torch.float32)
        x2 = x.reshape(-1, self.shape)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        out = torch.empty_like(x2, memory_format=torch.contiguous_format)
        _cabi.check(
            _cabi.load().phc_running_norm_forward(
                x2.data_ptr(), x2.shape[0], self.shape, x2.stride(0), self.running_mean.data_ptr(),
                self.running_var.data_ptr(), self.epsilon, self.clip, out.data_ptr(), out.stride(0),
                _cabi.stream_ptr(x.device),
            ),
            "phc_running_norm_forward",
        )  # fmt: skip
        return out.view(x.shape)

    @torch.no_grad()
    def moments(self, x: torch.Tensor, sums: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Accumulate the per-column fp64 [sum | sum of squares] of x [rows, shape] into ``sums``."""
        _cabi.require_cuda(x, "x", torch.float32)
        assert x.dim() == 2 and x.shape[1] == self.shape, "x must be 2D [rows, shape]"
        if x.stride(-1) != 1:
            x = x.contiguous()
        if sums is None:
            sums = torch.zeros(2 * self.shape, dtype=torch.float64, device=x.device)
        _cabi.check(
            _cabi.load().phc_obs_moments(
                x.data_ptr(), x.shape[0], self.shape, x.stride(0), sums.data_ptr(), _cabi.stream_ptr(x.device)
            ),
            "phc_obs_moments",
        )
        return sums

    def enable_peer_reduce(self, group=None, timeout_ms: int = 0):
        """Multi-GPU: do the all-reduce and the blend in ONE kernel over NVLink peer memory instead of
        NCCL all-reduce + update kernel (``parallel.PeerReduce``).  Collective: every rank of ``group`` calls it."""
        from .parallel import PeerReduce

        self._peers = PeerReduce.from_process_group(self.shape, self.running_mean.device, group, timeout_ms)
        return self._peers

    @torch.no_grad()
    def update_from_moments(self, sums: torch.Tensor, rows: int, group=None):
        """Blend statistics given local partials; all-reduces them first when distributed.  With
        ``enable_peer_reduce`` the reduction and the blend are one launch and ``sums`` is zeroed by it."""
        if getattr(self, "_peers", None) is not None:
            self._peers.update(self.running_mean, self.running_var, self.count, sums, rows)
            return
        n = 2 * self.shape
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            payload = reduce_moments(sums, rows, group)
        else:
            # one process: no exchange.  The payload lives in a persistent buffer filled by device-side copies, so the
            # whole update can be captured in a CUDA graph together with the steps that produced the partials.
            payload = getattr(self, "_payload", None)
            if payload is None or payload.device != sums.device:
                payload = self._payload = torch.empty(n + 1, dtype=torch.float64, device=sums.device)
            payload[:n].copy_(sums)
            if isinstance(rows, torch.Tensor):
                payload[n:].copy_(rows.reshape(1))
            else:
                payload[n:].fill_(float(rows))
        _cabi.check(
            _cabi.load().phc_running_norm_update(
                self.running_mean.data_ptr(), self.running_var.data_ptr(), self.count.data_ptr(),
                payload.data_ptr(), payload[n:].data_ptr(), self.shape, _cabi.stream_ptr(sums.device),
            ),
            "phc_running_norm_update",
        )  # fmt: skip

    @torch.no_grad()
    def update(self, x: torch.Tensor, group=None):  # :23-34
        x = x.float()
        self.update_from_moments(self.moments(x), x.shape[0], group)
