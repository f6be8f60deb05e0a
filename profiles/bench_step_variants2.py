"""Round 2: what each optional epilogue adds to the T = 1 fused step (us per launch, 128-launch CUDA graph over a ring of
buffer sets larger than L2, best of 5): RunningNorm moments as TMA bulk reductions and as atomics, the fused
normaliser, the episode bookkeeping, the reset of the flagged envs inside the step.

    python profiles/bench_step_variants2.py [num_envs]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, RunningNorm, _cabi, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
K = 128
R = max(4, -(-320 * (1 << 20) // (N * (312 + 934) * 4)))
capi = _cabi.load()
print(f"# step kernel variants at N = {N} (us per launch, {K}-launch CUDA graph, ring of {R} buffer sets, best of 5)\n")
print("| variant | us |\n|---|---|")
VARIANTS = [
    ("plain", {}),
    ("+ moments, one atomic per sum (default)", dict(moments=2)),
    ("+ moments, TMA bulk reductions (PHC_OPT_MOMENTS_BULK)", dict(moments=1)),
    ("+ normaliser fp32", dict(norm=torch.float32)),
    ("+ normaliser bf16", dict(norm=torch.bfloat16)),
    ("+ episode bookkeeping", dict(ep=True)),
    ("+ bookkeeping + reset in the step (synthetic state: ~25 % flagged per step)", dict(ep=True, reset=True)),
    ("+ bookkeeping + reset in the step, termination out of reach (clip ends only)", dict(ep=True, reset=True, far=True)),
    ("+ bookkeeping + reset + moments + normaliser bf16 (everything)", dict(ep=True, reset=True, far=True, moments=2, norm=torch.bfloat16)),
]
if os.environ.get("TERM_DISTS"):  # the in-step reset at intermediate flag rates: termination distances in metres
    VARIANTS = [("+ episode bookkeeping", dict(ep=True))] + [
        (f"+ bookkeeping + reset in the step, termination distance {d} m", dict(ep=True, reset=True, term=float(d)))
        for d in os.environ["TERM_DISTS"].split(",")]
for name, v in VARIANTS:
    capi.phc_set_option(_cabi.OPT_MOMENTS_BULK, 1 if v.get("moments") == 1 else 0)
    envs, states = [], []
    first = None
    for r in range(R):
        env = HumanoidPHC(lib, N, device=dev, obs_moments=bool(v.get("moments")) and r == 0)
        ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
        env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
        if first is None:
            env.set_clock(clock)
            first = env
        else:
            for k in ("progress_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset", "_sampled_motion_ids",
                      "_obs_moment_buckets"):
                setattr(env, k, getattr(first, k))
        if v.get("far"):
            env.set_termination_distances(torch.full((24,), 1e6, device=dev))
        if v.get("term"):
            env.set_termination_distances(torch.full((24,), v["term"], device=dev))
        if v.get("norm") is not None:
            env.set_obs_normalizer(RunningNorm(934, device=dev), dtype=v["norm"])
        if v.get("ep"):
            pe = PHCPufferEnv(env, log_interval=1 << 30, fused=False)
            env.set_episode_buffers(dict(terminals=pe.terminals, truncations=pe.truncations, masks=pe.masks,
                                         episode_returns=pe.episode_returns, episode_lengths=pe.episode_lengths,
                                         sums=torch.zeros((1024, _cabi.EPISODE_SUM_COLS), dtype=torch.float64, device=dev)))
            env._pe = pe
        if v.get("reset"):
            env.enable_auto_reset(True)
        envs.append(env)
    prog0 = first.progress_buf.clone()
    start0, goff0 = first._motion_start_times.clone(), first._global_offset.clone()

    def run(k):
        for i in range(k):
            if i % R == 0:
                first.progress_buf.copy_(prog0)
                if v.get("reset"):  # the resets move the clock: put it back with the progress
                    first._motion_start_times.copy_(start0)
                    first._global_offset.copy_(goff0)
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
    extra = ""
    if v.get("reset"):
        extra = f" ({100 * float(envs[(K - 1) % R].extras.get('reset', envs[(K - 1) % R]._reset_out).float().mean()):.1f} % flagged by the last step)"
    print(f"| {name}{extra} | {best:.2f} |", flush=True)
    del envs, g
    torch.cuda.empty_cache()
capi.phc_set_option(_cabi.OPT_MOMENTS_BULK, 0)
