"""What an integrator WITHOUT CUDA graphs sees: wall-clock per eager call of the shim's public API (Python + ctypes +
launch), 4096 envs, one synchronize at the end of 2000 calls; then a cProfile of the same loop.

    python profiles/bench_env_eager.py [num_envs]
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
CALLS = 2000
print(f"# eager calls at N = {N}: wall-clock us per call over {CALLS} calls (one synchronize at the end)\n")
print("| call | us / call |\n|---|---|")


def timed(fn, calls=CALLS):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    return t_cpu / calls * 1e6, t_all / calls * 1e6


def make(fused, **kw):
    env = HumanoidPHC(lib, N, device=dev, use_power_reward=True, **kw)
    env.set_termination_distances(torch.full((24,), 1e6, device=dev))  # resets at clip ends only
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
    env.set_sim_state(synth.make_sim_state(ref, seed=1236))
    env.set_clock(clock)
    return env, PHCPufferEnv(env, log_interval=1 << 30, fused=fused)


actions = torch.rand(N, 69, device=dev) * 2.4 - 1.2
phase = torch.rand(N, device=dev)
env, penv = make(True)
cpu, tot = timed(lambda: env.post_physics_step(True))
print(f"| HumanoidPHC.post_physics_step (one launch) | {cpu:.1f} submit, {tot:.1f} with the final synchronize |")
cpu, tot = timed(lambda: env.step(actions))
print(f"| HumanoidPHC.step(actions) (two launches) | {cpu:.1f} submit, {tot:.1f} |")
cpu, tot = timed(lambda: penv.step(actions, phase))
print(f"| PHCPufferEnv(fused=True).step (two launches: bookkeeping, reward copy, reset inside) | {cpu:.1f} submit, {tot:.1f} |")
env2, penv2 = make(False)
cpu, tot = timed(lambda: penv2.step(actions, phase))
print(f"| PHCPufferEnv(fused=False).step (round-1 launches) | {cpu:.1f} submit, {tot:.1f} |")

pr = cProfile.Profile()
pr.enable()
for _ in range(CALLS):
    penv.step(actions, phase)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print("\n```\n" + "\n".join(line for line in s.getvalue().splitlines() if line.strip())[:3500] + "\n```")
