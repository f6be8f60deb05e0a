"""Where a PHCPufferEnv.step goes: cumulative prefixes of the loop (fused bookkeeping, resets at clip ends only), each
prefix as a 128-step CUDA graph.  The difference between consecutive rows is what a piece costs IN the loop (its own
time plus the dependent-launch gap), as opposed to alone (bench_env_breakdown.py).

    python profiles/bench_env_prefix.py [num_envs]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
K = 128
env = HumanoidPHC(lib, N, device=dev, use_power_reward=True)
env.set_termination_distances(torch.full((24,), 1e6, device=dev))
ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
env.set_sim_state(synth.make_sim_state(ref, seed=1236))
env.set_clock(clock)
penv = PHCPufferEnv(env, log_interval=1 << 30, fused=True)
actions = torch.rand(N, 69, device=dev) * 2.4 - 1.2
phase = torch.rand(N, device=dev)
state0 = env._rigid_body_state_reshaped.clone()

pieces = [
    ("state write-back stand-in (copy_ of 5 MB)", lambda: env._rigid_body_state_reshaped.copy_(state0)),
    ("+ clamp actions", lambda: torch.clamp(actions, -1, 1, out=penv.actions)),
    ("+ fused step (power reward, episode bookkeeping)", lambda: env.step(penv.actions)),
    ("+ reward clone", lambda: penv.rewards.clone()),
    ("+ reset_done", lambda: env.reset_done(phase)),
]


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            g.replay()
            e1.record(s)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    return best


print(f"# PHCPufferEnv.step prefixes at N = {N} (us per step, {K}-step CUDA graph, best of 5)\n")
print("| loop up to and including | us / step | this piece in the loop |\n|---|---|---|")
prev = 0.0
for i in range(len(pieces)):
    def prefix(i=i):
        for _, f in pieces[: i + 1]:
            f()
    env.set_clock(clock)
    t = timed(prefix)
    print(f"| {pieces[i][0]} | {t:.2f} | {t - prev:.2f} |", flush=True)
    prev = t
env.set_clock(clock)
t = timed(lambda: (env._rigid_body_state_reshaped.copy_(state0), penv.step(actions, phase)))
print(f"| PHCPufferEnv.step itself | {t:.2f} | |")
# the reset alone on the flags of the last step of the loop above
m = (penv.terminals | penv.truncations).clone()
t = timed(lambda: env._reset_masked(m, phase))
print(f"| phc_reset_envs alone, {100 * float(m.float().mean()):.1f} % of the envs flagged | {t:.2f} | |")
# the same without the stand-in copy and with the step alone
env.set_clock(clock)
t = timed(lambda: env.step(penv.actions))
print(f"| fused step alone, back to back | {t:.2f} | |")
