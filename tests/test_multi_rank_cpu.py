"""CPU, world_size 2, gloo: the host logic of the multi-GPU path — env partitioning and the
RunningNorm statistics all-reduce — gives every rank the statistics of the concatenated batch."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from humanoid_b200.parallel import env_partition, reduce_moments
from oracle import phc_oracle as O


def test_env_partition_covers_everything_once():
    for n in (1, 7, 8, 4096, 4099, 65536, 65541):
        for w in (1, 2, 3, 4, 8):
            spans = [env_partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            for lo, hi in spans[:-1]:
                assert (lo % 8 == 0 or lo == n) and (hi % 8 == 0 or hi == n)
    assert env_partition(65536, 8, 3) == (24576, 32768)
    with pytest.raises(ValueError):
        env_partition(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(n_total, 934, generator=g) * 2.5 + 0.7  # every rank draws the same batch
        lo, hi = env_partition(n_total, world, rank)
        xs = x[lo:hi].double()
        sums = torch.cat([xs.sum(0), (xs * xs).sum(0)])
        payload = reduce_moments(sums, hi - lo)
        n = payload[-1]
        mean = (payload[:934] / n).float()
        var = (payload[934:1868] / n - (payload[:934] / n) ** 2).float()
        m0, v0, c0 = torch.zeros(1, 934), torch.ones(1, 934), torch.ones(1)
        want_m, want_v, _ = O.running_norm_update(m0, v0, c0, x)  # single-process update on the whole batch
        w = 1 / c0
        got_m = m0 * (1 - w) + mean * w
        got_v = v0 * (1 - w) + var * w
        ret[rank] = (
            float(n),
            float((got_m - want_m).abs().max()),
            float(((got_v - want_v).abs() / want_v.abs()).max()),
            payload.clone(),
        )
    finally:
        dist.destroy_process_group()


def test_running_norm_allreduce_world2_gloo():
    world, n_total = 2, 1000
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == n_total and r1[0] == n_total
    assert torch.equal(r0[3], r1[3]), "ranks disagree after the all-reduce"
    for r in (r0, r1):
        assert r[1] < 1e-5 and r[2] < 1e-5  # within 1e-5 of the single-process update


def test_reduce_moments_without_process_group_is_identity():
    s = torch.arange(6, dtype=torch.float64)
    p = reduce_moments(s, 3)
    assert p.tolist() == [0, 1, 2, 3, 4, 5, 3]
    with pytest.raises(TypeError):
        reduce_moments(s.float(), 3)
