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


def test_clip_sampler_matches_reference_sampling_state(golden):
    """Host logic of the loader (which clips, PMCP weights, batch sampling) against the reference's MotionLibSMPL
    under the same torch seed: load_data ordering, load_motions :300-321, update_*_sampling_weight, sample_motions."""
    import types

    import numpy as np

    from humanoid_b200.motion_build import ClipSampler

    _, _, ref_ml = ref_loader.load()
    from puffer_phc.poselib_skeleton import SkeletonTree

    g = golden("motion_build")
    nf = g.inp("num_frames").tolist()
    starts = np.concatenate([[0], np.cumsum(nf)])
    clips = {}
    for m in range(len(nf)):
        sl = slice(int(starts[m]), int(starts[m + 1]))
        clips[f"clip{m}"] = {"root_trans_offset": g.inp("root_trans_offset")[sl].clone(), "pose_aa": g.inp("pose_aa")[sl].numpy().copy(),
                             "pose_quat_global": g.inp("pose_quat_global")[sl].numpy().copy(), "beta": np.zeros(16), "fps": int(g.inp("fps")[m])}  # fmt: skip
    names = [f"b{j}" for j in range(24)]
    N = 10
    trees = [SkeletonTree(names, g.inp("parent_indices").long(), g.inp("local_translation")[e % len(nf)]) for e in range(N)]
    betas = [torch.zeros(17) for _ in range(N)]
    limbs = [np.zeros(10) for _ in range(N)]

    for im_eval in (False, True):
        ref = object.__new__(ref_ml.MotionLibSMPL)
        ref.m_cfg = types.SimpleNamespace(max_length=-1, fix_height=ref_ml.FixHeightMode.no_fix, is_deterministic=False,
                                          im_eval=im_eval, num_thread=1)  # fmt: skip
        ref._device, ref.mesh_parsers = "cpu", None
        # load_data (:192-230) minus joblib.load of a file
        order = sorted(clips.items(), key=lambda e: len(e[1]["pose_quat_global"]), reverse=True) if im_eval else clips.items()
        ref._motion_data_list = np.array([v for _, v in order])
        ref._motion_data_keys = np.array([k for k, _ in order])
        ref._num_unique_motions = len(clips)
        ref.setup_constants(fix_height=ref_ml.FixHeightMode.no_fix, num_thread=1)
        mine = ClipSampler()
        mine._init_clips(clips, im_eval=im_eval)
        assert mine._motion_data_keys.tolist() == ref._motion_data_keys.tolist()

        failed = ["clip3", "clip0"]
        for step in range(3):
            if step == 1:
                ref.update_soft_sampling_weight(failed), mine.update_soft_sampling_weight(failed)
            if step == 2:
                ref.update_hard_sampling_weight(failed[:1]), mine.update_hard_sampling_weight(failed[:1])
            assert torch.equal(mine._sampling_prob, ref._sampling_prob) and torch.equal(mine._termination_history, ref._termination_history)
            torch.manual_seed(100 + step)
            np.random.seed(0)
            ref.load_motions(trees, betas, limbs, random_sample=True)
            torch.manual_seed(100 + step)
            ids = mine.pick_clips(N, True, 0, None, is_deterministic=False)
            assert torch.equal(ids, ref._curr_motion_ids) and mine.curr_motion_keys.tolist() == ref.curr_motion_keys.tolist()
            assert torch.equal(mine._sampling_batch_prob, ref._sampling_batch_prob)
            torch.manual_seed(7)
            a = ref.sample_motions(5)
            torch.manual_seed(7)
            assert torch.equal(mine.sample_motions(5), a)
        # the sequential branch with a start index (evaluation sweep, humanoid_phc.py:1381-1402)
        ref.load_motions(trees, betas, limbs, random_sample=False, start_idx=4)
        assert torch.equal(mine.pick_clips(N, False, 4, None, is_deterministic=False), ref._curr_motion_ids)
        ref.update_soft_sampling_weight([]), mine.update_soft_sampling_weight([])
        assert torch.equal(mine._sampling_prob, ref._sampling_prob)


def test_reference_step_orchestration_equals_the_oracle_step_bit_for_bit():
    """``ref_loader.reference_step`` — what ``bench.py --impl reference`` times (the reference's own functions in the
    env's order) — against ``oracle.step`` on the same inputs, T = 1 and T = 3."""
    for T in (1, 3):
        lib_data, clock, state = synth.make_case(query=_ref_query, num_envs=300, num_motions=24, seed=5 + T, max_progress=20)
        term = torch.full((24,), 0.25)
        p1, p2 = clock.progress_buf.clone(), clock.progress_buf.clone()
        want = O.step(O.OracleMotionLib(lib_data), state, p1, clock.motion_start_times, clock.motion_start_times_offset,
                      clock.global_offset, clock.sampled_motion_ids, term, synth.SIM_DT, time_steps=T)  # fmt: skip
        got = ref_loader.reference_step(ref_loader.make_reference_lib(lib_data), state, p2, clock.motion_start_times,
                                        clock.motion_start_times_offset, clock.global_offset, clock.sampled_motion_ids,
                                        term, synth.SIM_DT, time_steps=T)  # fmt: skip
        assert torch.equal(p1, p2)
        for a, b, nm in zip(got, want, ("obs", "reward", "reward_raw", "reset", "terminated")):
            assert_equal_exact(a, b, f"T={T}: {nm}")
