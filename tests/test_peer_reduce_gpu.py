"""GPU: the fused all-reduce + RunningNorm blend over peer memory (phc_running_norm_update_peers).

One-device tests run W "ranks" of one process on W streams (``connect_local``): the protocol — epoch flags, two
alternating slots, rank-ordered sums, timeout — is the same code as across GPUs.  The CUDA-IPC path needs two
devices: ``test_two_processes_over_cuda_ipc`` launches torchrun and is skipped on a one-GPU box."""

import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, assert_close

pytestmark = pytest.mark.gpu

C = 934


def _ranks(world, timeout_ms=0):
    from humanoid_b200 import RunningNorm
    from humanoid_b200.parallel import PeerReduce

    peers = [PeerReduce(r, world, C, "cuda", timeout_ms) for r in range(world)]
    PeerReduce.connect_local(peers)
    rns = [RunningNorm(C, device="cuda") for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    return peers, rns, streams


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_peer_update_equals_single_process_update_on_the_concatenated_batch(world):
    from oracle import phc_oracle as O

    peers, rns, streams = _ranks(world)
    gen = torch.Generator().manual_seed(world)
    m, v, c = torch.zeros(1, C), torch.ones(1, C), torch.ones(1)
    sums = [torch.zeros(2 * C, dtype=torch.float64, device="cuda") for _ in range(world)]
    for rollout in range(5):  # both slots, several epochs, uneven row counts (one rank may hold nothing)
        rows = [int(torch.randint(0 if world > 1 else 1, 300, (1,), generator=gen)) for _ in range(world)]
        rows[0] = max(rows[0], 1)
        xs = [torch.randn(n, C, generator=gen) * (1 + rollout) + 0.3 * rollout for n in rows]
        torch.cuda.synchronize()
        for r in reversed(range(world)):  # launch order must not matter
            with torch.cuda.stream(streams[r]):
                if rows[r]:
                    rns[r].moments(xs[r].cuda(), sums[r])
                peers[r].update(rns[r].running_mean, rns[r].running_var, rns[r].count, sums[r], rows[r])
        torch.cuda.synchronize()
        m, v, c = O.running_norm_update(m, v, c, torch.cat(xs))  # running_norm.py:23-34 on the whole batch
        for r in range(world):
            assert peers[r].status() == rollout + 1
            assert_close(rns[r].running_mean.cpu(), m, rtol=1e-5, atol=1e-6, what=f"mean, rank {r}")
            assert_close(rns[r].running_var.cpu(), v, rtol=1e-5, atol=1e-6, what=f"var, rank {r}")
            assert float(rns[r].count) == float(c) and float(sums[r].abs().max()) == 0.0
            assert torch.equal(rns[r].running_mean, rns[0].running_mean), "ranks must agree bit for bit"
            assert torch.equal(rns[r].running_var, rns[0].running_var)


def test_peer_update_matches_the_two_kernel_path_and_is_graph_capturable():
    from humanoid_b200 import RunningNorm

    world = 2
    peers, rns, streams = _ranks(world)
    gen = torch.Generator().manual_seed(9)
    xs = [(torch.randn(257 + 100 * r, C, generator=gen) * 3 - 1).cuda() for r in range(world)]
    plain = RunningNorm(C, device="cuda")
    total = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    for x in xs:
        plain.moments(x, total)
    plain.update_from_moments(total, sum(x.shape[0] for x in xs))  # no process group: sums are already global
    sums = [rns[r].moments(xs[r]) for r in range(world)]
    keep = [s.clone() for s in sums]
    graphs = []
    torch.cuda.synchronize()
    for r in range(world):  # capture one launch per rank, replay them concurrently
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[r]):
            peers[r].update(rns[r].running_mean, rns[r].running_var, rns[r].count, sums[r], xs[r].shape[0])
        graphs.append(g)
    for rep in range(3):
        for r in range(world):
            sums[r].copy_(keep[r])
            rns[r].running_mean.zero_(), rns[r].running_var.fill_(1.0), rns[r].count.fill_(1.0)
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        for r in range(world):
            assert peers[r].status() == rep + 1
            assert_close(rns[r].running_mean, plain.running_mean, rtol=1e-6, atol=1e-7, what="mean vs two-kernel path")
            assert_close(rns[r].running_var, plain.running_var, rtol=1e-6, atol=1e-7, what="var vs two-kernel path")


def test_missing_peer_times_out_instead_of_hanging():
    from humanoid_b200 import _cabi

    peers, rns, _ = _ranks(2, timeout_ms=50)
    sums = torch.ones(2 * C, dtype=torch.float64, device="cuda")
    before = rns[0].running_mean.clone()
    peers[0].update(rns[0].running_mean, rns[0].running_var, rns[0].count, sums, 10)  # rank 1 never launches
    with pytest.raises(_cabi.PhcError, match="peer"):
        peers[0].status()
    assert torch.equal(rns[0].running_mean, before) and float(rns[0].count) == 1.0 and float(sums.min()) == 1.0
    assert peers[0].status() == 0  # the status word is cleared once reported; nothing completed


def test_timeout_then_resync_then_retry_gives_the_right_statistics():
    """A timed-out launch decides once for all its blocks (nothing is half-updated), leaves the ranks out of step,
    and phc_peer_reduce_resync puts the group back: the retried update equals the oracle on both ranks."""
    from humanoid_b200 import _cabi
    from oracle import phc_oracle as O

    peers, rns, streams = _ranks(2, timeout_ms=50)
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(100 + 57 * r, C, generator=gen) * 2 + 0.5 for r in range(2)]
    sums = [rns[r].moments(xs[r].cuda()) for r in range(2)]
    keep0 = sums[0].clone()
    # one good epoch first, so that the failure happens at a non-zero epoch with both slots' history in place
    for r in range(2):
        with torch.cuda.stream(streams[r]):
            peers[r].update(rns[r].running_mean, rns[r].running_var, rns[r].count, sums[r].clone(), xs[r].shape[0])
    torch.cuda.synchronize()
    m, v, c = O.running_norm_update(torch.zeros(1, C), torch.ones(1, C), torch.ones(1), torch.cat(xs))
    before = (rns[0].running_mean.clone(), rns[0].running_var.clone(), float(rns[0].count))
    peers[0].update(rns[0].running_mean, rns[0].running_var, rns[0].count, sums[0], xs[0].shape[0])  # rank 1 absent
    with pytest.raises(_cabi.PhcError, match="peer"):
        peers[0].status()
    assert torch.equal(rns[0].running_mean, before[0]) and torch.equal(rns[0].running_var, before[1])
    assert float(rns[0].count) == before[2] and torch.equal(sums[0], keep0), "a timed-out launch must touch nothing"
    peers[0].resync(local=peers)
    for r in reversed(range(2)):
        with torch.cuda.stream(streams[r]):
            peers[r].update(rns[r].running_mean, rns[r].running_var, rns[r].count, sums[r], xs[r].shape[0])
    torch.cuda.synchronize()
    m, v, c = O.running_norm_update(m, v, c, torch.cat(xs))
    for r in range(2):
        assert peers[r].status() == 1  # epochs restart after a resync
        assert_close(rns[r].running_mean.cpu(), m, rtol=1e-5, atol=1e-6, what=f"mean after retry, rank {r}")
        assert_close(rns[r].running_var.cpu(), v, rtol=1e-5, atol=1e-6, what=f"var after retry, rank {r}")
        assert float(rns[r].count) == float(c) and float(sums[r].abs().max()) == 0.0
    assert torch.equal(rns[0].running_mean, rns[1].running_mean) and torch.equal(rns[0].running_var, rns[1].running_var)


def test_bad_arguments():
    from humanoid_b200 import _cabi
    from humanoid_b200.parallel import PeerReduce

    with pytest.raises(_cabi.PhcError):
        PeerReduce(2, 2, C, "cuda")
    with pytest.raises(_cabi.PhcError):
        PeerReduce(0, 17, C, "cuda")
    p = PeerReduce(0, 2, C, "cuda")
    rn_m, rn_v, cnt = torch.zeros(1, C, device="cuda"), torch.ones(1, C, device="cuda"), torch.ones(1, device="cuda")
    with pytest.raises(_cabi.PhcError):  # not connected yet
        p.update(rn_m, rn_v, cnt, torch.zeros(2 * C, dtype=torch.float64, device="cuda"), 1)
    with pytest.raises(_cabi.PhcError):
        p.connect([b"x" * 64])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_processes_over_cuda_ipc():
    """torchrun, one rank per GPU: IPC handles exchanged through the process group, mailboxes read over NVLink;
    against the NCCL all-reduce path and the single-process update."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29731", os.path.join(ROOT, "tests", "peer_reduce_worker.py")],
        env=env, capture_output=True, text=True, timeout=300,
    )  # fmt: skip
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PEER_REDUCE_OK" in out.stdout


def test_rollout_cycle_step_epilogue_to_fused_update():
    """The whole statistics path of a rollout on two 'ranks': the step's moments epilogue accumulates while stepping,
    take_obs_moments() folds the buckets, ONE fused launch per rank exchanges and blends — against
    RunningNorm.update(experience.obs) of the reference (policies/running_norm.py:23-34, scripts/phc_train.py:331-332)
    on the concatenated rows of both ranks."""
    from humanoid_b200 import HumanoidPHC, MotionLib, synth
    from oracle import phc_oracle as O

    world, steps = 2, 5
    peers, rns, streams = _ranks(world)
    envs, rows = [], []
    for r in range(world):
        N = 1000 + 37 * r
        lib_data = synth.make_motion_lib(32, 60, 120, (30,), seed=40 + r, device="cuda")
        lib = MotionLib(lib_data, device="cuda")
        clock = synth.make_clock(lib_data, N, seed=50 + r, max_progress=20)
        ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
        env = HumanoidPHC(lib, N, device="cuda", obs_moments=True)
        env.set_sim_state(synth.make_sim_state(ref, seed=60 + r))
        env.set_clock(clock)
        envs.append(env)
    m, v, c = torch.zeros(1, C), torch.ones(1, C), torch.ones(1)
    for rollout in range(2):
        experience = []
        for _ in range(steps):
            for env in envs:
                env.step()
                experience.append(env.obs_buf.cpu().clone())
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                streams[r].wait_stream(torch.cuda.current_stream())
                sums, n = envs[r].take_obs_moments()
                assert n == steps * envs[r].num_envs
                peers[r].update(rns[r].running_mean, rns[r].running_var, rns[r].count, sums, n)
        torch.cuda.synchronize()
        m, v, c = O.running_norm_update(m, v, c, torch.cat(experience))
        for r in range(world):
            assert_close(rns[r].running_mean.cpu(), m, rtol=1e-5, atol=1e-6, what=f"mean, rollout {rollout}, rank {r}")
            assert_close(rns[r].running_var.cpu(), v, rtol=1e-5, atol=1e-6, what=f"var, rollout {rollout}, rank {r}")
        assert torch.equal(rns[0].running_mean, rns[1].running_mean)
