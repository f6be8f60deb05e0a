"""Small driver for ncu: builds the configs[1] workload (4096 envs, one clip per env) and
launches the fused step kernel a few dozen times outside any CUDA graph.

    python profiles/prof_step.py [num_envs] [time_steps] [launches] [config4]

PROF_MOMENTS=1: with the RunningNorm moments (obs_moments); PROF_RESET=1: with the episode bookkeeping and the reset
of the flagged envs inside the step (termination distance out of reach: clip ends only).

With a 4th argument ``config4`` the workload is BASELINE configs[3] instead: a 10k-clip mixed-fps
library, random motion ids and unaligned times, a different clock for every ring slot.
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, _cabi, synth  # noqa: E402

MOM = os.environ.get("PROF_MOMENTS") == "1"
RST = os.environ.get("PROF_RESET") == "1"

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 40
C4 = len(sys.argv) > 4 and sys.argv[4] == "config4"
dev = torch.device("cuda", 0)
if C4:
    g = torch.Generator(device=dev).manual_seed(1241)
    nf = torch.randint(100, 701, (10000,), generator=g, device=dev)
    nf[:20] = 7000
    lib_data = synth.make_motion_lib(10000, fps_choices=(30, 60, 120), seed=1234, device=dev, frames_per_motion=nf)
else:
    lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
R = max(4, -(-320 * (1 << 20) // (N * (312 + 358 + 576 * T) * 4)))  # ring of buffer sets larger than the 126 MB L2
envs = []
for r in range(R):
    if C4:
        clock = synth.make_clock(lib_data, N, seed=1235 + 31 * r, ids="random", aligned=False, max_progress=30)
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1 if C4 else r + 1), clock.global_offset)
    env = HumanoidPHC(lib, N, device=dev, time_steps=T, obs_moments=MOM and r == 0)
    env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
    env.set_clock(clock)
    if MOM and r > 0:
        env._obs_moment_buckets = envs[0]._obs_moment_buckets
    if RST:
        env.set_termination_distances(torch.full((24,), 1e6, device=dev))
        pe = PHCPufferEnv(env, log_interval=1 << 30, fused=False)
        env.set_episode_buffers(dict(terminals=pe.terminals, truncations=pe.truncations, masks=pe.masks,
                                     episode_returns=pe.episode_returns, episode_lengths=pe.episode_lengths,
                                     sums=torch.zeros((1024, _cabi.EPISODE_SUM_COLS), dtype=torch.float64, device=dev)))
        env._pe = pe
        env.enable_auto_reset(True)
    envs.append(env)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
prog0 = [e.progress_buf.clone() for e in envs] if C4 else envs[0].progress_buf.clone()
start0, goff0 = envs[0]._motion_start_times.clone(), envs[0]._global_offset.clone()
if not C4:  # all ring slots share one motion clock, re-seeded at every lap like bench.py does
    for e in envs[1:]:
        for k in ("progress_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset", "_sampled_motion_ids"):
            setattr(e, k, getattr(envs[0], k))
for i in range(L):
    if i == L // 2:
        e0.record()
    if i % R == 0:
        if C4:
            for e, p0 in zip(envs, prog0):
                e.progress_buf.copy_(p0)
        else:
            envs[0].progress_buf.copy_(prog0)
            if RST:  # the resets move the clock: put it back with the progress
                envs[0]._motion_start_times.copy_(start0)
                envs[0]._global_offset.copy_(goff0)
    envs[i % R].post_physics_step(True)
e1.record()
torch.cuda.synchronize()
print(f"N={N} T={T}{' config4' if C4 else ''}: {1e3 * e0.elapsed_time(e1) / (L - L // 2):.2f} us per launch (back-to-back, no graph)")
