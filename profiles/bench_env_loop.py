"""Everything the env does around the physics step, through the shim's public API, in a CUDA graph.
fused=False: PHCPufferEnv.step = clamp actions -> PD targets -> HumanoidPHC.step (fused step + power reward [+ AMP
buffers]) -> reward clone -> episode bookkeeping -> device-side reset of the flagged envs (one launch each).
fused=True (round 2): TWO launches — action clip + PD targets; fused step with power reward, bookkeeping, reward copy
and the reset of the flagged envs inside it.

    python profiles/bench_env_loop.py [num_envs]
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
print(f"# Whole env step through the shim at N = {N} (us per step, {K}-step CUDA graph, best of 5)\n")
print("| variant | us / step | env-steps/s |\n|---|---|---|")
# the synthetic sim state terminates 24-31 % of the envs every step (it mixes both flag values for the parity tests);
# with the termination distance out of reach only the end of a clip resets an env — under 1 % of the envs per step,
# the rate of the real task (episodes of up to 300 steps)
for name, kw, fused, far in (("fused step only (HumanoidPHC.post_physics_step)", None, False, False),
                             ("PHCPufferEnv.step: power reward, episode bookkeeping, device-side resets", dict(use_power_reward=True), False, False),
                             ("PHCPufferEnv(fused=True): two launches (clip + PD targets; step with bookkeeping, reward copy, reset inside)", dict(use_power_reward=True), True, False),
                             ("the same + AMP observation buffers (10-step history)", dict(use_power_reward=True, use_amp_obs=True), True, False),
                             ("PHCPufferEnv(fused=False), resets at clip ends only", dict(use_power_reward=True), False, True),
                             ("PHCPufferEnv(fused=True), resets at clip ends only", dict(use_power_reward=True), True, True),
                             ("the same + AMP observation buffers", dict(use_power_reward=True, use_amp_obs=True), True, True)):
    env = HumanoidPHC(lib, N, device=dev, **(kw or {}))
    if far:
        env.set_termination_distances(torch.full((24,), 1e6, device=dev))
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
    env.set_sim_state(synth.make_sim_state(ref, seed=1236))
    env.set_clock(clock)
    penv = PHCPufferEnv(env, log_interval=1 << 30, use_amp_obs=bool(kw and kw.get("use_amp_obs")), fused=fused)
    actions = torch.rand(N, 69, device=dev) * 2.4 - 1.2
    phase = torch.rand(N, device=dev)
    state0 = env._rigid_body_state_reshaped.clone()

    def one():
        if kw is None:
            env.post_physics_step(True)
        else:
            env._rigid_body_state_reshaped.copy_(state0)  # stands in for the physics write-back
            penv.step(actions, phase)

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                one()
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
    rate = "" if kw is None else f" ({100 * float((penv.terminals | penv.truncations).float().mean()):.1f} % of the envs reset by the last step)"
    print(f"| {name}{rate} | {best:.2f} | {N / best * 1e6:.3g} |", flush=True)
