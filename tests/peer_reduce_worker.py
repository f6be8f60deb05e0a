"""Worker of tests/test_peer_reduce_gpu.py::test_two_processes_over_cuda_ipc (launched by torchrun, one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humanoid_b200 import RunningNorm  # noqa: E402
from oracle import phc_oracle as O  # noqa: E402

C = 934
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
gen = torch.Generator().manual_seed(3)
fused, nccl = RunningNorm(C, device="cuda"), RunningNorm(C, device="cuda")
fused.enable_peer_reduce(timeout_ms=20000)
m, v, c = torch.zeros(1, C), torch.ones(1, C), torch.ones(1)
for rollout in range(4):
    xs = [torch.randn(200 + 50 * r + rollout, C, generator=gen) * (1 + r) for r in range(world)]  # same draws on every rank
    x = xs[rank].cuda()
    s = fused.moments(x)
    fused.update_from_moments(s, x.shape[0])  # ONE launch: P2P all-reduce + blend
    nccl.update(x)  # moments + NCCL all-reduce + blend
    m, v, c = O.running_norm_update(m, v, c, torch.cat(xs))
    torch.cuda.synchronize()
    assert fused._peers.status() == rollout + 1
    for name, got, want in (("mean", fused.running_mean, m), ("var", fused.running_var, v)):
        excess = float(((got.cpu() - want).abs() - (1e-6 + 1e-5 * want.abs())).max())  # rtol 1e-5, atol 1e-6
        assert excess <= 0, (name, excess)
        other = getattr(nccl, "running_" + name)
        excess = float(((got - other).abs() - (1e-7 + 1e-6 * other.abs())).max())
        assert excess <= 0, ("vs nccl", name, excess)
    assert float(s.abs().max()) == 0.0
    # bit-identical on every rank
    both = [torch.empty_like(fused.running_mean) for _ in range(world)]
    dist.all_gather(both, fused.running_mean)
    assert all(torch.equal(b, both[0]) for b in both)
# latency of the fused launch against all-reduce + update kernel
s = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
for which, fn in (("fused", lambda: fused.update_from_moments(s, 100)), ("nccl", lambda: nccl.update_from_moments(s, 100))):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"PEER_REDUCE_TIMING {which} {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per update (world {world})")
dist.barrier()
if rank == 0:
    print("PEER_REDUCE_OK")
dist.destroy_process_group()
