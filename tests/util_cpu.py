"""Helpers shared by the CPU tests."""

import torch

from humanoid_b200 import synth
from oracle import phc_oracle as O


def oracle_env_from_golden(g, use_amp_obs=True):
    lib = O.OracleMotionLib(synth.MotionData(**g.group("in.lib")))
    c = synth.Clock(**g.group("in.clock"))
    if use_amp_obs:
        amp = dict(dof_subset=g.inp("dof_subset"), key_body_ids=g.inp("key_body_ids"),
                   num_amp_obs_steps=int(g.inp("num_amp_obs_steps")))  # fmt: skip
    else:  # the reset-mode fixture runs without the AMP buffers
        amp = dict(dof_subset=torch.arange(69), key_body_ids=torch.zeros(0, dtype=torch.long), num_amp_obs_steps=2,
                   use_amp_obs=False)  # fmt: skip
    env = O.OracleEnv(
        lib, c.progress_buf.shape[0], c.progress_buf, c.motion_start_times, c.motion_start_times_offset,
        c.global_offset, c.sampled_motion_ids, g.inp("pd_action_offset"), g.inp("pd_action_scale"),
        rew_power_coef=float(g.inp("rew_power_coef")), termination_distance=float(g.inp("termination_distance")), **amp,
    )  # fmt: skip

    def write_sim(state, dof_state, dof_force):
        env.state[:], env.dof_state[:], env.dof_force[:] = state, dof_state, dof_force
        env.root_states[:] = state[:, 0]

    env.write_sim = write_sim
    env.read_sim = lambda: (env.state, env.root_states, env.dof_state)
    return env
