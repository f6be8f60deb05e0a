"""Fused RunningNorm.forward epilogue vs step + standalone forward kernel (config 2, CUDA graph).

    python profiles/bench_norm.py [num_envs]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, RunningNorm, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
rn = RunningNorm(934, device=dev)
R, K = 17, 256


def make(fused, moments=False):
    envs = []
    for r in range(R):
        ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
        env = HumanoidPHC(lib, N, device=dev, obs_moments=moments)
        if moments and r:
            env._obs_moment_buckets = envs[0]._obs_moment_buckets  # one accumulator for the rollout
        env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
        env.set_clock(clock)
        if r:
            env.progress_buf = envs[0].progress_buf
        if fused:
            env.set_obs_normalizer(rn, dtype=fused)
        envs.append(env)
    return envs


def timed(envs, after):
    p0 = envs[0].progress_buf.clone()

    def run(k):
        for i in range(k):
            if i % R == 0:
                envs[0].progress_buf.copy_(p0)
            envs[i % R].post_physics_step(True)
            after(envs[i % R])

    run(R)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            run(K)
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


outs = [torch.empty(N, 934, device=dev) for _ in range(R)]
plain = timed(make(False), lambda e: None)
mom = timed(make(False, moments=True), lambda e: None)
sums = torch.zeros(2 * 934, dtype=torch.float64, device=dev)
mom_sep = timed(make(False), lambda e: rn.moments(e.obs_buf, sums))
sep = timed(make(False), lambda e: rn(e.obs_buf))
fused = timed(make(torch.float32), lambda e: None)
fused16 = timed(make(torch.bfloat16), lambda e: None)
print(f"# RunningNorm.forward at N = {N} (us per step, {K}-step CUDA graph, best of 5)\n")
print("| variant | us / step |\n|---|---|")
print(f"| step only (raw obs) | {plain:.2f} |")
print(f"| step with the RunningNorm moments epilogue (obs_moments: fp64 column sums accumulated in the step, 32 buckets) | {mom:.2f} |")
print(f"| step + standalone phc_obs_moments over the same rows | {mom_sep:.2f} |")
print(f"| step + standalone phc_running_norm_forward (allocates its output) | {sep:.2f} |")
print(f"| step with the fused epilogue (raw + normalised fp32 rows) | {fused:.2f} |")
print(f"| step with the fused epilogue, normalised rows as bf16 (PHC_STEP_OBS_NORM_BF16) | {fused16:.2f} |")
