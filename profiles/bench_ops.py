"""Per-function kernels (K1-K5) and the fused step: time and achieved algorithmic GB/s.

Algorithmic bytes per row from SURVEY §8(d): get_motion_state 6508 B/query (all 13 outputs),
self obs 2680 B, imitation obs v6 4800 B, reward 2516 B, reset 581 B, fused step 8804 B.

    python profiles/bench_ops.py [num_envs] [out.md]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import (HumanoidPHC, MotionLib, compute_humanoid_im_reset, compute_humanoid_observations_smpl_max,  # noqa: E402
                           compute_imitation_observations_v6, compute_imitation_reward, synth)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
dev = torch.device("cuda", 0)
M = min(N, 8192)
lib_data = synth.make_motion_lib(M, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
t = synth.reward_time(clock, extra_steps=1)
ref = lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
state = synth.make_sim_state(ref, seed=1236)
pos, rot, vel, ang = synth.body_views(state)
env = HumanoidPHC(lib, N, device=dev)
env.set_sim_state(state)
env.set_clock(clock)
prog = (clock.progress_buf + 1).to(torch.int16)
pass_time = t >= lib_data.motion_lengths[clock.sampled_motion_ids]
term = torch.full((24,), 0.25, device=dev)
reset_in = torch.ones(N, dtype=torch.bool, device=dev)
rwd = env.rwd_specs

OPS = {
    "K1 get_motion_state (13 outputs)": (6508, lambda: lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)),
    "K2 self obs smpl_max": (2680, lambda: compute_humanoid_observations_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)),
    "K3 imitation obs v6 (T=1)": (4800, lambda: compute_imitation_observations_v6(pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], 1, True)),
    "K4 imitation reward": (2516, lambda: compute_imitation_reward(pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], rwd)),
    "K5 im reset": (581, lambda: compute_humanoid_im_reset(reset_in, prog, None, None, pos, ref["rg_pos"], pass_time, True, term, False)),
    "K6 fused step": (8804, lambda: env.post_physics_step(False)),
    "per-function step (2xK1 + K4 + K5 + K2 + K3 + cat)": (8804, lambda: (env._compute_reward(), env._compute_reset(), env._compute_observations())),
}  # fmt: skip
print(f"# Per-function kernels at N = {N}\n", file=out)
print("`eager` = a Python call of the drop-in (ctypes wrapper + output allocation + launch, CUDA events around 30 calls);\n"
      "`graph` = the same 20 calls captured in a CUDA graph and replayed (device time only — what the kernel itself costs).\n"
      "GB/s and the fraction are for the graph column.\n", file=out)
print("| op | eager us / call | graph us / call | algorithmic B / row | achieved GB/s | of measured 6456 GB/s |\n|---|---|---|---|---|---|", file=out)
for name, (bytes_row, fn) in OPS.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    # device time: capture K calls in a graph (outputs come from the graph's private pool)
    K = 20
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for _ in range(K):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        with torch.cuda.stream(stream):
            e0.record(stream)
            g.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, 1e3 * e0.elapsed_time(e1) / K)
    del g
    gbs = bytes_row * N / best / 1e3
    print(f"| {name} | {us:.1f} | {best:.1f} | {bytes_row} | {gbs:.0f} | {gbs / 6456.2:.2f} |", file=out)
