"""Round 2: the persistent warp-specialised step kernel (K6-persist) against K6-fast at multi-wave sizes
(us per step, 64-step CUDA graph over a ring of buffer sets larger than L2, best of 5; algorithmic fraction of the
measured HBM peak at 8804 B per env-step).

    python profiles/bench_persist.py [sizes...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, _cabi, synth  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 32768, 65536]
dev = torch.device("cuda", 0)
capi = _cabi.load()
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6456.2
K = 64
print(f"# K6-persist vs K6-fast (us per step, {K}-step graph, best of 5; frac of {peak} GB/s at 8804 B/env-step)\n")
print("| envs | K6-fast us | frac | K6-persist (3 slots, 5 blocks/SM) us | frac | K6-persist (4 slots, 4 blocks/SM) us | frac |\n|---|---|---|---|---|---|---|")
for N in sizes:
    lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
    R = max(4, -(-320 * (1 << 20) // (N * (312 + 934) * 4)))
    envs, first = [], None
    for r in range(R):
        env = HumanoidPHC(lib, N, device=dev)
        ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=r + 1), clock.global_offset)
        env.set_sim_state(synth.make_sim_state(ref, seed=1236 + r), copy=False)
        del ref
        if first is None:
            env.set_clock(clock)
            first = env
        else:
            for k in ("progress_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset", "_sampled_motion_ids"):
                setattr(env, k, getattr(first, k))
        envs.append(env)
    prog0 = first.progress_buf.clone()

    def run(k):
        for i in range(k):
            if i % R == 0:
                first.progress_buf.copy_(prog0)
            envs[i % R].post_physics_step(True)

    row = []
    for mode in (0, 2, 3):
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
        del g
    capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    f = lambda us: 8804 * N / (us * 1e-6) / 1e9 / peak  # noqa: E731
    print(f"| {N} | {row[0]:.2f} | {f(row[0]):.3f} | {row[1]:.2f} | {f(row[1]):.3f} | {row[2]:.2f} | {f(row[2]):.3f} |", flush=True)
    del envs, first, lib, lib_data
    torch.cuda.empty_cache()
