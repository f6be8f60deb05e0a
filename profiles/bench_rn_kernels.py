import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from humanoid_b200 import RunningNorm
rn = RunningNorm(934, device="cuda")
for rows in (4096, 32768, 131072):
    x = torch.randn(rows, 934, device="cuda")
    sums = torch.zeros(2*934, dtype=torch.float64, device="cuda")
    for _ in range(3): rn.moments(x, sums)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): rn.moments(x, sums)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/20*1e3
    print(f"phc_obs_moments rows={rows}: {us:.1f} us, {rows*934*4/us/1e3:.0f} GB/s")
    out = rn(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): out = rn(x)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1)/20*1e3
    print(f"phc_running_norm_forward rows={rows}: {us:.1f} us, {2*rows*934*4/us/1e3:.0f} GB/s")
