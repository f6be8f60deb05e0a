"""The fused step kernel's instantiations back to back (128 launches in a CUDA graph, same sim state every step):
what the power reward and the fused episode bookkeeping add to the plain step.

    python profiles/bench_step_variants.py [num_envs]
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
print(f"# step kernel variants at N = {N} (us per launch, {K}-launch CUDA graph, best of 5)\n")
print("| variant | us |\n|---|---|")
for name, power, fused in (("plain", False, False), ("+ power reward", True, False), ("+ episode bookkeeping", False, True),
                           ("+ power reward + episode bookkeeping", True, True)):
    env = HumanoidPHC(lib, N, device=dev, use_power_reward=power)
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
    env.set_sim_state(synth.make_sim_state(ref, seed=1236))
    env.set_clock(clock)
    penv = PHCPufferEnv(env, log_interval=1 << 30, fused=fused)
    for _ in range(3):
        env.post_physics_step(True)
    env.set_clock(clock)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                env.post_physics_step(True)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(6):
        env.set_clock(clock)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s)
            g.replay()
            e1.record(s)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K * 1e3)
    print(f"| {name} | {best:.2f} |", flush=True)
