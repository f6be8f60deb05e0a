"""Small driver for ncu: builds the configs[1] workload (4096 envs, one clip per env) and
launches the fused step kernel a few dozen times outside any CUDA graph.

    python profiles/prof_step.py [num_envs] [time_steps] [launches]
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
R = 17
envs = []
for r in range(R):
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
    env = HumanoidPHC(lib, N, device=dev, time_steps=T)
    env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
    env.set_clock(clock)
    envs.append(env)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(L):
    if i == L // 2:
        e0.record()
    envs[i % R].post_physics_step(True)
e1.record()
torch.cuda.synchronize()
print(f"N={N} T={T}: {1e3 * e0.elapsed_time(e1) / (L - L // 2):.2f} us per launch (back-to-back, no graph)")
