"""Round 2: a 32-step rollout with the RunningNorm moments accumulated by the step (config 3 sizes), us per step:
plain step / step with the moments epilogue, for the kernel phc_step_fused picks and with the persistent kernel off.

    python profiles/bench_rollout_moments.py [sizes...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, _cabi, synth  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [8192, 16384, 32768, 65536]
dev = torch.device("cuda", 0)
capi = _cabi.load()
K = 32
print(f"# rollout of {K} steps, us per step (graph, ring larger than L2, best of 5)\n")
print("| envs | plain (default kernel) | + moments, own rows (default) | + moments, atomics | K6-fast plain | K6-fast + moments own rows | K6-fast + moments atomics |\n|---|---|---|---|---|---|---|")
for N in sizes:
    lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
    R = max(4, -(-320 * (1 << 20) // (N * (312 + 934) * 4)))
    row = []
    for persist in (1, 0):
        for mom, own in ((False, 1), (True, 1), (True, 0)):
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, persist)
            capi.phc_set_option(_cabi.OPT_MOMENTS_OWN, own)
            envs, first = [], None
            for r in range(R):
                env = HumanoidPHC(lib, N, device=dev, obs_moments=mom and r == 0)
                ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
                env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
                del ref
                if first is None:
                    env.set_clock(clock)
                    first = env
                else:
                    for k in ("progress_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset",
                              "_sampled_motion_ids", "_obs_moment_buckets"):
                        setattr(env, k, getattr(first, k))
                envs.append(env)
            prog0 = first.progress_buf.clone()

            def run(k):
                for i in range(k):
                    if i % R == 0:
                        first.progress_buf.copy_(prog0)
                    envs[i % R].post_physics_step(True)

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
            del g, envs, first
            torch.cuda.empty_cache()
    capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    capi.phc_set_option(_cabi.OPT_MOMENTS_OWN, 1)
    print(f"| {N} | " + " | ".join(f"{x:.2f}" for x in row) + " |", flush=True)
    del lib, lib_data
    torch.cuda.empty_cache()
