"""Round 2, session 2: the step with the RunningNorm moments (obs_moments) on K6-fast (1868 L2 adds per four envs) and
on the persistent kernel with the partials in registers (1868 adds per block, at the end), next to the plain step
(us per step, 64-step CUDA graph over a ring of buffer sets larger than L2, best of 5).

    python profiles/bench_moments_sizes.py [sizes...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, _cabi, synth  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 32768, 65536]
dev = torch.device("cuda", 0)
capi = _cabi.load()
K = 64
print(f"# step with obs_moments: K6-fast (atomics per block of four envs) vs K6-persist (partials in registers) — us per step, {K}-step graph, best of 5\n")
print("| envs | plain, default kernel | K6-fast + moments | K6-persist + moments | persistent / fast |\n|---|---|---|---|---|")
for N in sizes:
    lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
    R = max(4, -(-320 * (1 << 20) // (N * (312 + 934) * 4)))
    row = []
    for moments, mode in ((False, 1), (True, 0), (True, 2)):
        envs, first = [], None
        for r in range(R):
            env = HumanoidPHC(lib, N, device=dev, obs_moments=moments and r == 0)
            ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
            env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
            del ref
            if first is None:
                env.set_clock(clock)
                first = env
            else:
                for k in ("progress_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset", "_sampled_motion_ids",
                          "_obs_moment_buckets"):
                    setattr(env, k, getattr(first, k))
            envs.append(env)
        prog0 = first.progress_buf.clone()

        def run(k):
            for i in range(k):
                if i % R == 0:
                    first.progress_buf.copy_(prog0)
                envs[i % R].post_physics_step(True)

        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, mode)
        run(R)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
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
        row.append(best)
        del envs, g, first
        torch.cuda.empty_cache()
    capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    print(f"| {N} | {row[0]:.2f} | {row[1]:.2f} | {row[2]:.2f} | {row[2] / row[1]:.2f} |", flush=True)
    del lib, lib_data
    torch.cuda.empty_cache()
