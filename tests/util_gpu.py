"""Helpers shared by the GPU parity tests: oracle <-> CUDA plumbing."""

import torch

from humanoid_b200 import HumanoidPHC, MotionLib, synth
from oracle import phc_oracle as O

DEV = "cuda"

# north_star tolerance: 1e-5 relative (fp32) for observations and rewards, with the 1e-6 absolute term SURVEY §8(c)
# suggests for entries that cancel to ~0.  "Relative" is taken against the NORM OF THE VECTOR an entry is a component
# of (conftest.natural_scale: a body's 3-vector, a tangent-normal 6-vector, a quaternion): the dominant error is the
# reference's own 1-ulp atan2f in the heading, which rotates a body vector v by ~1e-7 rad and so moves each of its
# components by ~1e-7 |v| whatever the component's size.  profiles/r2_parity_margins.md has the measured margins.
OBS_TOL = dict(rtol=1e-5, atol=1e-6, scale="vec")
# dof_pos = angle * xyz / sqrt(1 - w*w) of a slerp output: the 1e-5 bar holds for every joint whose rotation is not
# tiny; for near-identity joints (|w| > 0.99, i.e. angle < 0.283 rad) the reference's own sqrt(1 - w*w) amplifies a
# 1-ulp difference in w (libdevice vs SLEEF sin / acos) to ~3e-5 relative, and only those joints get 1e-4.
DOF_TOL = dict(rtol=1e-5, atol=1e-6, scale="vec", small_angle=dict(below=0.283, rtol=1e-4, atol=2e-5))
# tensors that CONTAIN such joint angles among other things (the interleaved dof state, AMP observation rows built
# from exp_map_to_quat of dof_pos): the small-angle tolerance for the whole tensor
DOF_DERIVED_TOL = dict(rtol=1e-4, atol=2e-5)


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
