#!/usr/bin/env python
"""Benchmark of the PHC step path: env-steps/s for the full post-physics step
(motion query @t + reward + reset + motion query @t+dt + self obs + imitation obs v6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU torch path

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
synthetic sim state: BASELINE.json configs[1] — num_envs = 4096 per GPU, one 60-300 frame
30 fps clip per env (ids == arange, the reference regime), frame-aligned start times.
Multi-GPU: envs and their clips are partitioned across ranks (weak scaling, no collective on
the step); the RunningNorm all-reduce is timed separately and reported under "rms".

`value`  : device-resident inputs; K steps captured in ONE CUDA graph (one kernel per step),
           timed with CUDA events on the launching stream, max over ranks.  Every step reads a
           different sim-state buffer and writes a different obs buffer out of a ring larger
           than the 126 MB L2, so inputs come from HBM and outputs go to HBM.
`e2e`    : the same metric through the C-ABI host pipeline (phc_host_step) with pinned HOST
           buffers: H2D of the sim state + clock, the fused kernel, D2H of obs/reward/flags,
           all inside the timed region.
`cpu_baseline` / `--impl reference`: oracle/phc_oracle.py — the torch-CPU restatement that
           matches the reference bit for bit — on the box's host cores (kind "port"; the
           reference is Python and cannot travel to the GPU box).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

NUM_ENVS_PER_GPU = 4096
BYTES_PER_ENV_STEP = 8804  # SURVEY §8(d) / BASELINE.md §3: T=1, dt-aligned (3 distinct frames)
BYTES_PER_ENV_STEP_T = lambda T: 1248 + 54 + (T + 2) * 1248 + (358 + 576 * T) * 4 + 22  # noqa: E731
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def workload_name(n, workload="config2", clips=0, frames=0):
    if workload == "config4":
        return (f"PHC SMPL 24-body step compute, num_envs={n} per GPU, AMASS-scale synthetic library ({clips} clips, "
                f"{frames} frames, mixed 30/60/120 fps), random motion ids and unaligned times per step")
    return (f"PHC SMPL 24-body step compute (motion query + smpl_max obs + imitation obs v6 + reward + reset), "
            f"num_envs={n} per GPU, one 60-300 frame 30 fps synthetic clip per env, synthetic sim state")  # fmt: skip


# ---------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------
# CPU baseline: the oracle port on host cores
# ---------------------------------------------------------------------------------------
def cpu_oracle_rate(lib_data, clock, state, steps, warmup, threads):
    """env-steps/s of oracle/phc_oracle.py::step (torch CPU fp32) on CPU copies of the workload."""
    from humanoid_b200 import synth
    from oracle import phc_oracle as O

    torch.set_num_threads(threads)
    lib = O.OracleMotionLib(lib_data)
    term = torch.full((24,), 0.25)
    n = state.shape[0]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            prog = clock.progress_buf.clone()
            t0 = time.perf_counter()
            O.step(lib, state, prog, clock.motion_start_times, clock.motion_start_times_offset,
                   clock.global_offset, clock.sampled_motion_ids, term, synth.SIM_DT)  # fmt: skip
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
    total = sum(times)
    return n * len(times) / total, 1e3 * total / len(times)


def run_reference(args):
    """--impl reference: the reference's CPU torch path (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from humanoid_b200 import synth
    from oracle import phc_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = NUM_ENVS_PER_GPU

    def query(lib_data, ids, times, offset):
        return O.OracleMotionLib(lib_data).get_motion_state(ids, times, offset)

    # a bounded sample: when K is large, shrink the env batch so K steps stay within ~2 minutes
    lib_data, clock, state = synth.make_case(n, n, query, seed=1234, device="cpu", max_progress=30)
    _, ms = cpu_oracle_rate(lib_data, clock, state, steps=2, warmup=1, threads=threads)
    budget_ms = 120e3
    n_run = n
    if ms * (args.steps + args.warmup) > budget_ms:
        frac = budget_ms / (ms * (args.steps + args.warmup))
        n_run = max(256, int(n * frac) // 256 * 256)
    sl = slice(0, n_run)
    clock_s = synth.Clock(**{k: v[sl] for k, v in clock.__dict__.items()})
    rate, ms = cpu_oracle_rate(lib_data, clock_s, state[sl], steps=args.steps, warmup=args.warmup, threads=threads)
    sample = f"{n_run} of {n} envs per step x {args.steps} steps, torch CPU fp32 oracle port, {threads} threads"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n), "num_envs_per_gpu": n, "time_steps": 1, "sample_envs": n_run},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    emit(line)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(T):
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        return d.get(f"step_kernel_T{T}_N{NUM_ENVS_PER_GPU}")
    except Exception:
        return None


def run_gpu(args):
    import torch.distributed as dist

    from humanoid_b200 import HumanoidPHC, MotionLib, _cabi, synth
    from humanoid_b200.running_norm import RunningNorm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    N, T, K, W = args.num_envs, args.time_steps, args.steps, args.warmup
    W = max(W, 3)
    seed = 1234 + 1000 * rank  # each rank owns its own envs and their clips

    # -------- workload (device-generated; the reference pose comes from the product's own query kernel)
    obs_dim = 358 + 576 * T
    set_bytes = N * (24 * 13 + obs_dim) * 4
    R = max(4, -(-args.ring_mb * (1 << 20) // set_bytes))  # ring of R buffer sets > L2
    envs = []
    config4 = args.workload == "config4"
    if config4:
        # BASELINE configs[3]: AMASS-scale library (~10k clips, ~4M frames, mixed fps, a few very long
        # clips) replicated per GPU; every ring slot has its OWN clock with random motion ids and
        # unaligned start times, so consecutive steps gather unrelated frames (defeats L2).
        M = args.clips
        g = torch.Generator(device=dev).manual_seed(seed + 7)
        nf = torch.randint(100, 701, (M,), generator=g, device=dev)
        nf[: max(1, M // 500)] = 7000
        lib_data = synth.make_motion_lib(M, fps_choices=(30, 60, 120), seed=seed, device=dev, frames_per_motion=nf)
    else:
        lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=seed, device=dev)
    lib = MotionLib(lib_data, device=dev)
    clock = None
    first = None
    for r in range(R):
        if config4 or clock is None:
            clock = synth.make_clock(lib_data, N, seed=seed + 1 + 31 * r, ids="random" if config4 else "mod",
                                     aligned=not config4, max_progress=30)  # fmt: skip
        t = synth.reward_time(clock, extra_steps=1 if config4 else r + 1)
        ref = lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
        state = synth.make_sim_state(ref, seed=seed + 2 + r)
        env = HumanoidPHC(lib, N, device=dev, time_steps=T)
        env.set_sim_state(state, copy=False)
        if first is None or config4:
            env.set_clock(clock)
            first = first or env
        else:  # all ring slots share one motion clock
            env.progress_buf = first.progress_buf
            env._motion_start_times = first._motion_start_times
            env._motion_start_times_offset = first._motion_start_times_offset
            env._global_offset = first._global_offset
            env._sampled_motion_ids = first._sampled_motion_ids
        envs.append(env)
    progress0 = envs[0].progress_buf.clone()
    prog0_all = [e.progress_buf.clone() for e in envs] if config4 else None
    torch.cuda.synchronize()

    def run_steps(k):
        for i in range(k):
            if config4:
                if i >= R and i % R == 0:  # every lap: each slot's clock back to its start
                    for e, p0 in zip(envs, prog0_all):
                        e.progress_buf.copy_(p0)
            elif i % R == 0:
                first.progress_buf.copy_(progress0)  # re-seed the clock at every lap of the ring
            envs[i % R].post_physics_step(True)

    # warm-up (also sets the kernel's smem attribute outside capture)
    run_steps(max(W, R))  # every ring slot at least once
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            run_steps(K)
    torch.cuda.synchronize()
    graph.replay()  # one untimed replay
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    reps = args.repeats
    ms_runs = []
    for _ in range(reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()  # EXACTLY K steps
        e1.record()
        barrier()
        ms_runs.append(e0.elapsed_time(e1))
    if rank == 0:
        time.sleep(0.2)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = statistics.median(ms_runs)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = N * world * K / (ms_total * 1e-3)
    ms_per_step = ms_total / K
    frac_term = float(envs[(K - 1) % R]._terminate_buf.float().mean().item())

    # -------- e2e: host buffers through the C-ABI pipeline
    import ctypes as C

    def pinned(t):
        return t.detach().cpu().contiguous().pin_memory()

    e0 = envs[0]  # the host-buffer leg and the CPU baseline run ring slot 0's inputs
    clock = synth.Clock(progress_buf=progress0, motion_start_times=e0._motion_start_times,
                        motion_start_times_offset=e0._motion_start_times_offset, global_offset=e0._global_offset,
                        sampled_motion_ids=e0._sampled_motion_ids)  # fmt: skip
    h_state = pinned(envs[0]._rigid_body_state_reshaped)
    h_prog0 = pinned(progress0)
    h_prog = h_prog0.clone().pin_memory()
    h_start, h_off = pinned(clock.motion_start_times), pinned(clock.motion_start_times_offset)
    h_goff, h_ids = pinned(clock.global_offset), pinned(clock.sampled_motion_ids)
    h_obs = torch.empty((N, obs_dim), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
    h_raw = torch.empty((N, 4), dtype=torch.float32).pin_memory()
    h_reset = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_term = torch.empty(N, dtype=torch.uint8).pin_memory()
    capi = _cabi.load()
    ctx = C.c_void_p()
    term_host = (C.c_float * 24)(*([0.25] * 24))
    spec = _cabi.reward_spec(first.rwd_specs)
    _cabi.check(capi.phc_host_step_create(lib.handle, N, T, args.e2e_chunks, term_host, 0xFFFFFF, 0, 1,
                                          synth.SIM_DT, C.byref(spec), C.byref(ctx)), "phc_host_step_create")  # fmt: skip
    hargs = _cabi.PhcHostStepArgs(
        h_state.data_ptr(), h_prog.data_ptr(), h_start.data_ptr(), h_off.data_ptr(), h_goff.data_ptr(),
        h_ids.data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(), h_raw.data_ptr(), h_reset.data_ptr(), h_term.data_ptr(),
    )  # fmt: skip
    Ke = min(K, args.e2e_steps)
    for _ in range(3):
        h_prog.copy_(h_prog0)
        _cabi.check(capi.phc_host_step(ctx, C.byref(hargs), N), "phc_host_step")
    # the host path must agree with the device path on the same inputs
    envs[0].progress_buf.copy_(progress0)
    envs[0].post_physics_step(True)
    torch.cuda.synchronize()
    e2e_ok = bool(torch.equal(envs[0].obs_buf.cpu(), h_obs) and torch.equal(envs[0].rew_buf.cpu(), h_rew))
    e2e_runs = []
    for _ in range(3):  # median of three timed loops of Ke calls (each call returns with the outputs in host memory)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            if i % R == 0:
                h_prog.copy_(h_prog0)
            _cabi.check(capi.phc_host_step(ctx, C.byref(hargs), N), "phc_host_step")
        t1 = time.perf_counter()
        barrier()
        e2e_s = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(e2e_s.item()))
    e2e_value = N * world * Ke / statistics.median(e2e_runs)
    h2d = int(capi.phc_host_step_h2d_bytes(ctx, N))
    d2h = int(capi.phc_host_step_d2h_bytes(ctx, N))
    capi.phc_host_step_destroy(ctx)

    # -------- RunningNorm: moments of a 32-step rollout + ONE exchange (+ blend), per rollout.
    # The exchange + blend is one launch per rank over NVLink peer memory (phc_running_norm_update_peers);
    # the NCCL all-reduce + update-kernel path is timed beside it.
    rn = RunningNorm(obs_dim, device=dev)
    rn_nccl = RunningNorm(obs_dim, device=dev)
    peer_note = None
    if world > 1:
        try:
            rn.enable_peer_reduce(timeout_ms=20000)
            ok = torch.ones(1, device=dev)
        except Exception as exc:  # e.g. no P2P between two of the devices: keep the NCCL path, say so in the line
            peer_note = f"peer path unavailable on rank {rank}: {exc}"
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks or none
        if float(ok.item()) == 0.0:
            rn._peers = None
            peer_note = peer_note or "peer path unavailable on another rank"
    sums = torch.zeros(2 * obs_dim, dtype=torch.float64, device=dev)
    roll = min(32, R)

    def rms_once():
        sums.zero_()
        for r in range(roll):
            rn.moments(envs[r].obs_buf, sums)
        rn.update_from_moments(sums, roll * N)

    def timed_ms(fn, reps=1):
        fn()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            fn()
        r1.record()
        barrier()
        t = torch.tensor([r0.elapsed_time(r1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rms_ms = timed_ms(rms_once)
    exch_fused_us = timed_ms(lambda: rn.update_from_moments(sums, roll * N), reps=50) * 1e3
    exch_nccl_us = timed_ms(lambda: rn_nccl.update_from_moments(sums, roll * N), reps=50) * 1e3
    if world > 1 and rn._peers is not None:
        try:
            rn._peers.status()  # raises if any launch timed out waiting for a peer
        except Exception as exc:
            peer_note = f"peer launch failed: {exc}"
    fused_on = world > 1 and getattr(rn, "_peers", None) is not None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peak()
    bytes_step = BYTES_PER_ENV_STEP if T == 1 else BYTES_PER_ENV_STEP_T(T)
    if config4 and T == 1:
        bytes_step = 10052  # 4 distinct frames
    achieved = bytes_step * N / (ms_per_step * 1e-3) / 1e9  # per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(N, args.workload, lib_data.num_motions, lib_data.total_frames),
            "num_envs_per_gpu": N, "num_envs_total": N * world, "time_steps": T,
            "motion_clips_per_gpu": lib_data.num_motions, "motion_frames_per_gpu": lib_data.total_frames,
            "l2": f"inputs larger than L2: every step uses the next of {R} sim-state/obs buffer sets "
                  f"({R * set_bytes / 2**20:.0f} MiB ring), clock re-seeded each lap",
            "cuda_graph": f"{K} step kernels in one graph", "terminated_frac": round(frac_term, 4),
            "timing_repeats": reps, "seed": 1234,
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "steps": Ke, "timing_repeats": 3, "chunks": args.e2e_chunks, "matches_device_path": e2e_ok},
        "gpu_launches": K,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(T), "peak_source": peak_src, "bytes_per_env_step": bytes_step,
                     "kernel": "phc::step_fast_kernel<4,8,false,false>" if T == 1
                     else "phc::step_multi2_kernel", "launch_ms": ms_per_step},
        "rms": {"what": f"RunningNorm.update over a {roll}-step rollout: fp64 column moments + "
                        + (f"one fused launch per rank (all-reduce of {(2 * obs_dim + 1) * 8} B over NVLink peer memory + blend)"
                           if fused_on else f"NCCL all-reduce of {(2 * obs_dim + 1) * 8} B + blend" if world > 1
                           else "blend (1 GPU: no exchange)"),
                "ms_per_rollout": rms_ms,
                "exchange_and_blend_us": {"fused_peer_kernel" if fused_on else "update_path": exch_fused_us,
                                          "nccl_all_reduce_plus_update_kernel" if world > 1 else "update_path_again": exch_nccl_us},
                **({"note": peer_note} if peer_note else {})},
    }  # fmt: skip

    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        lib_cpu = lib_data.to("cpu")
        clock_cpu = synth.Clock(**{k: v.cpu() for k, v in clock.__dict__.items()})
        clock_cpu.progress_buf = progress0.cpu()
        state_cpu = envs[0]._rigid_body_state_reshaped.cpu()
        _, ms1 = cpu_oracle_rate(lib_cpu, clock_cpu, state_cpu, steps=1, warmup=1, threads=threads)
        cs = max(5, min(100, int(15e3 / ms1)))  # about 15 s of CPU work
        rate, msc = cpu_oracle_rate(lib_cpu, clock_cpu, state_cpu, steps=cs, warmup=3, threads=threads)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": msc,
            "sample": f"{cs} steps of the same {N}-env workload (same tensors copied to host), "
                      f"oracle/phc_oracle.py torch CPU fp32, {threads} threads",
        }  # fmt: skip
    if world == 1 and args.torch_cuda_baseline:
        # the reference's own deployment: the same torch ops, eager, on this GPU (about 1100 launches per step)
        from oracle import phc_oracle as O

        lib_o = O.OracleMotionLib(lib_data)
        term = torch.full((24,), 0.25, device=dev)
        st = envs[0]._rigid_body_state_reshaped

        def torch_step():
            prog = progress0.clone()
            O.step(lib_o, st, prog, clock.motion_start_times, clock.motion_start_times_offset, clock.global_offset,
                   clock.sampled_motion_ids, term, synth.SIM_DT)  # fmt: skip

        with torch.no_grad():
            for _ in range(5):
                torch_step()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 30
            t0.record()
            for _ in range(reps):
                torch_step()
            t1.record()
            torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        line["torch_cuda_baseline"] = {"value": N / ms * 1e3, "unit": UNIT, "ms_per_step": ms,
                                       "what": "oracle/phc_oracle.py (the reference's torch ops, fp32) eager on this GPU"}  # fmt: skip
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def keep_stdout_for_the_json_line():
    """The contract is ONE line on stdout.  Libraries write there too (NCCL prints its version banner from C when
    NCCL_DEBUG is set in the environment), so fd 1 is pointed at stderr for the whole run and the JSON line goes to
    the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--num-envs", type=int, default=NUM_ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--time-steps", type=int, default=1, help="future reference frames T (config 5: 10)")
    ap.add_argument("--workload", choices=["config2", "config4"], default="config2",
                    help="config2: one clip per env, aligned (default, BASELINE configs[1]); config4: AMASS-scale "
                         "library, random ids / unaligned times per step (BASELINE configs[3])")
    ap.add_argument("--clips", type=int, default=10000, help="clips of the config4 library")
    ap.add_argument("--ring-mb", type=int, default=320, help="min size of the sim-state/obs buffer ring (> L2)")
    ap.add_argument("--repeats", type=int, default=5, help="timed graph replays; the median is reported")
    ap.add_argument("--e2e-steps", type=int, default=64)
    ap.add_argument("--e2e-chunks", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-cuda-baseline", action="store_true",
                    help="also time the reference's torch ops (oracle port) eager on this GPU; reported, not the target")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
