"""Stress the fused step for timing-dependent results (compute-sanitizer is not available on the pool).

    python profiles/stress_step.py [iters]

A. the same step replayed `iters` times with random filler work in between must give bit-identical outputs;
B. a memset of the output buffers immediately before the launch (the step kernel is launched with programmatic
   stream serialization: it must still be ordered after a non-kernel predecessor);
C. a pageable host->device copy of the sim state immediately before the launch;
D. back-to-back steps (PDL overlap) against the same steps separated by synchronisation;
E. (round 2, session 2) the step with the reset of the flagged envs inside it (spread over the block's warps by role)
   against step + reset_done(), three steps back to back, random filler work in between: every buffer bit-identical;
F. the persistent kernel with the RunningNorm partials in registers, three steps back to back: rows bit-identical to
   the plain step, partials equal to phc_obs_moments to 1e-12.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, RunningNorm, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
bad = 0
for N, norm_dtype in ((4099, torch.float32), (4096, torch.bfloat16), (1000, None), (8192, torch.float32)):
    lib_data = synth.make_motion_lib(64, 60, 200, (30,), seed=7, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=8, max_progress=30)
    ref = lib.get_motion_state(clock.sampled_motion_ids.to(dev), synth.reward_time(clock, extra_steps=1).to(dev), clock.global_offset.to(dev))
    state = synth.make_sim_state(ref, seed=9)
    state_host = state.cpu()
    env = HumanoidPHC(lib, N, device=dev, obs_moments=False)
    rn = RunningNorm(env.num_obs, device=dev)
    rn.running_mean.normal_()
    rn.running_var.uniform_(0.1, 2.0)
    if norm_dtype is not None:
        env.set_obs_normalizer(rn, dtype=norm_dtype)
    env.set_sim_state(state)
    env.set_clock(clock)
    env.step()
    torch.cuda.synchronize()
    outs = lambda: [t.clone() for t in [env.obs_buf, env.rew_buf, env.reward_raw, env.reset_buf, env._terminate_buf]
                    + ([env.obs_norm_buf] if norm_dtype is not None else [])]  # noqa: E731
    want = outs()
    gen = torch.Generator().manual_seed(1)
    for mode in "ABC":
        fails = 0
        for i in range(iters):
            env.set_clock(clock)
            if mode == "A":
                k = int(torch.randint(0, 4, (1,), generator=gen))
                for _ in range(k):
                    torch.empty(int(torch.randint(1, 1 << 22, (1,), generator=gen)), device=dev).normal_()
            elif mode == "B":
                env.obs_buf.zero_()
                env.rew_buf.zero_()
                if norm_dtype is not None:
                    env.obs_norm_buf.zero_()
            else:
                env._rigid_body_state_reshaped.zero_()
                env._rigid_body_state_reshaped.copy_(state_host)
            env.step()
            got = outs()
            if not all(torch.equal(a, b) for a, b in zip(got, want)):
                fails += 1
        print(f"N={N} norm={norm_dtype} mode {mode}: {fails}/{iters} iterations differ")
        bad += fails
    # D: three back-to-back steps vs the same with syncs
    def run3(sync):
        env.set_clock(clock)
        res = []
        for _ in range(3):
            env.step()
            if sync:
                torch.cuda.synchronize()
            res.append(env.obs_buf.clone())
        torch.cuda.synchronize()
        return res
    base = run3(True)
    fails = 0
    for i in range(iters // 4):
        fails += not all(torch.equal(a, b) for a, b in zip(run3(False), base))
    print(f"N={N} norm={norm_dtype} mode D: {fails}/{iters // 4} iterations differ")
    bad += fails
# ---- E: in-step reset ---------------------------------------------------------------------------------
from humanoid_b200 import _cabi  # noqa: E402

BUFS = ("obs_buf", "rew_buf", "reward_raw", "_rigid_body_state_reshaped", "_humanoid_root_states", "_dof_state",
        "progress_buf", "reset_buf", "_terminate_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset")
gen = torch.Generator().manual_seed(2)
for N, term in ((4096, 0.25), (4099, 0.6), (1000, 1e6)):
    lib_data = synth.make_motion_lib(N, 40, 90, (30,), seed=17, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=18, max_progress=45)
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
    state = synth.make_sim_state(ref, seed=19)
    a, b = HumanoidPHC(lib, N, device=dev), HumanoidPHC(lib, N, device=dev)
    b.enable_auto_reset(True)
    for e in (a, b):
        e.set_termination_distances(torch.full((24,), term, device=dev))
    fails = 0
    flagged = 0
    for i in range(iters // 4):
        for e in (a, b):
            e.set_sim_state(state)
            e.set_clock(clock)
        for k in range(3):
            phase = torch.rand(N, generator=gen).to(dev)
            a.step()
            a.reset_done(phase)
            for _ in range(int(torch.randint(0, 3, (1,), generator=gen))):
                torch.empty(int(torch.randint(1, 1 << 21, (1,), generator=gen)), device=dev).normal_()
            b.step(phase_by_env=phase)
            flagged += int(b.extras["reset"].sum())
            if not all(torch.equal(getattr(a, n), getattr(b, n)) for n in BUFS):
                fails += 1
    print(f"N={N} term={term} mode E: {fails}/{3 * (iters // 4)} steps differ ({flagged / (3 * (iters // 4) * N):.3f} of the envs flagged per step)")
    bad += fails
# ---- F: persistent kernel with the moments in registers ----------------------------------------------------
capi = _cabi.load()
for N in (8192, 16387, 3001):
    lib_data = synth.make_motion_lib(N, 60, 200, (30,), seed=27, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = synth.make_clock(lib_data, N, seed=28, max_progress=30)
    ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
    state = synth.make_sim_state(ref, seed=29)
    plain, env = HumanoidPHC(lib, N, device=dev), HumanoidPHC(lib, N, device=dev, obs_moments=True)
    rn = RunningNorm(934, device=dev)
    fails = 0
    for i in range(iters // 8):
        for e in (plain, env):
            e.set_sim_state(state)
            e.set_clock(clock)
        want = torch.zeros(2 * 934, dtype=torch.float64, device=dev)
        for k in range(3):
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 0)
            plain.step()
            rn.moments(plain.obs_buf, want)
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 2)
            env.step()
            if not (torch.equal(plain.obs_buf, env.obs_buf) and torch.equal(plain.rew_buf, env.rew_buf)
                    and torch.equal(plain.reset_buf, env.reset_buf)):
                fails += 1
        got, _ = env.take_obs_moments()
        if float(((got - want).abs() / want.abs().clamp_min(1.0)).max()) > 1e-12:
            fails += 1
    capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    print(f"N={N} mode F: {fails} failures in {iters // 8} rounds of three steps")
    bad += fails
print("TOTAL", bad)
sys.exit(1 if bad else 0)
