"""Helpers shared by the GPU parity tests: oracle <-> CUDA plumbing."""

import torch

from humanoid_b200 import HumanoidPHC, MotionLib, synth
from oracle import phc_oracle as O

DEV = "cuda"

# north_star tolerance: 1e-5 relative (fp32) for observations and rewards; the absolute term
# covers entries that cancel to ~0.
OBS_TOL = dict(rtol=1e-5, atol=2e-6)
# dof_pos = angle*axis/sin(theta) of a slerp output: for near-identity joint rotations the
# reference's own sqrt(1-w*w) amplifies a 1-ulp difference in w (libdevice vs SLEEF sin/acos)
# to ~3e-5 relative; see DESIGN.md "tolerances".
DOF_TOL = dict(rtol=1e-4, atol=2e-5)


def oracle_query(lib_data, ids, times, offset):
    return O.OracleMotionLib(lib_data).get_motion_state(ids, times, offset)


def make_case_cpu(**kw):
    return synth.make_case(query=oracle_query, device="cpu", **kw)


def env_from(lib_data, clock, state, time_steps=1, **kw):
    lib = MotionLib(lib_data, device=DEV)
    env = HumanoidPHC(lib, state.shape[0], device=DEV, bodies_per_env=state.shape[1], time_steps=time_steps, **kw)
    env.set_sim_state(state.to(DEV))
    env.set_clock(clock.to(DEV))
    return env


def oracle_step(lib_data, clock, state, time_steps=1, term=0.25, **kw):
    prog = clock.progress_buf.clone()
    out = O.step(
        O.OracleMotionLib(lib_data), state, prog, clock.motion_start_times, clock.motion_start_times_offset,
        clock.global_offset, clock.sampled_motion_ids, torch.full((24,), term), synth.SIM_DT,
        time_steps=time_steps, **kw,
    )  # fmt: skip
    return out + (prog,)


def clock_from_golden(g):
    c = g.group("in.clock")
    return synth.Clock(**c)


def lib_from_golden(g):
    return synth.MotionData(**g.group("in.lib"))
