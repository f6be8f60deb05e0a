"""Driver for ncu: a few eager PHCPufferEnv.step calls (fused bookkeeping, resets at clip ends only) at N = 4096.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python profiles/prof_env_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
env = HumanoidPHC(lib, N, device=dev, use_power_reward=True)
env.set_termination_distances(torch.full((24,), 1e6, device=dev))
ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
env.set_sim_state(synth.make_sim_state(ref, seed=1236))
env.set_clock(clock)
penv = PHCPufferEnv(env, log_interval=1 << 30, fused=True)
actions = torch.rand(N, 69, device=dev) * 2.4 - 1.2
phase = torch.rand(N, device=dev)
state0 = env._rigid_body_state_reshaped.clone()
for _ in range(STEPS):
    env._rigid_body_state_reshaped.copy_(state0)  # stands in for the physics write-back
    penv.step(actions, phase)
torch.cuda.synchronize()
print("resets in the last step: %.1f %%" % (100 * float((penv.terminals | penv.truncations).float().mean())))
