"""CPU, build container only: the oracle against the live reference on larger seeded inputs.

Skipped wherever /root/reference is absent (the GPU box); the committed fixtures under
tests/golden/ carry the same pin there."""

import pytest
import torch

from conftest import assert_equal_exact
from humanoid_b200 import synth
from oracle import phc_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _ref_query(lib_data, ids, times, offset):
    return ref_loader.make_reference_lib(lib_data).get_motion_state(ids, times, offset)


@pytest.mark.parametrize(
    "kw",
    [
        dict(num_envs=256, num_motions=16, seed=1, max_progress=20),  # BASELINE config 1
        dict(num_envs=512, num_motions=512, seed=2, max_frames=90, max_progress=20),  # ids == arange regime
        dict(num_envs=384, num_motions=40, seed=3, ids="random", aligned=False, fps_choices=(30, 60, 120),
             rot_regime="random", max_progress=30),
    ],
)  # fmt: skip
def test_whole_step_bit_exact_vs_reference(kw):
    _, ref_common, _ = ref_loader.load()
    lib_data, clock, state = synth.make_case(query=_ref_query, **kw)
    ref_lib = ref_loader.make_reference_lib(lib_data)
    lib = O.OracleMotionLib(lib_data)
    pos, rot, vel, ang = synth.body_views(state)
    dt = synth.SIM_DT
    term = torch.full((24,), 0.25)

    prog = clock.progress_buf.clone()
    obs, reward, raw, reset, terminated = O.step(
        lib, state, prog, clock.motion_start_times, clock.motion_start_times_offset, clock.global_offset,
        clock.sampled_motion_ids, term, dt,
    )  # fmt: skip

    p = clock.progress_buf.clone()
    p += 1
    t = p * dt + clock.motion_start_times + clock.motion_start_times_offset
    r0 = ref_lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    rr, rraw = ref_common.compute_imitation_reward(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang, r0["rg_pos"], r0["rb_rot"], r0["body_vel"], r0["body_ang_vel"],
        O.DEFAULT_RWD_SPECS,
    )  # fmt: skip
    rs, rt = ref_common.compute_humanoid_im_reset(
        torch.ones(state.shape[0], dtype=torch.bool), p, torch.zeros(1), torch.zeros(1), pos.clone(),
        r0["rg_pos"].clone(), t >= ref_lib._motion_lengths[clock.sampled_motion_ids], True, term, False,
    )  # fmt: skip
    t1 = (p + 1) * dt + clock.motion_start_times + clock.motion_start_times_offset
    r1 = ref_lib.get_motion_state(clock.sampled_motion_ids, t1, clock.global_offset)
    so = ref_common.compute_humanoid_observations_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    to = ref_common.compute_imitation_observations_v6(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang, r1["rg_pos"], r1["rb_rot"], r1["body_vel"], r1["body_ang_vel"], 1, True
    )  # fmt: skip
    assert torch.equal(obs, torch.cat([so, to], -1))
    assert torch.equal(reward, rr) and torch.equal(raw, rraw)
    assert_equal_exact(reset, rs, "reset")
    assert_equal_exact(terminated, rt, "terminated")
    assert 0 < terminated.float().mean() < 1


def test_motion_state_all_keys_bit_exact_vs_reference():
    lib_data = synth.make_motion_lib(24, 30, 120, fps_choices=(30, 60), seed=8)
    clock = synth.make_clock(lib_data, 300, seed=9, ids="random", aligned=False)
    t = synth.reward_time(clock) - 0.05  # some negative times
    ref = ref_loader.make_reference_lib(lib_data).get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    got = O.OracleMotionLib(lib_data).get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    for k, v in ref.items():
        assert torch.equal(got[k], v), k
