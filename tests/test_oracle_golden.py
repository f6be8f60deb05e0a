"""CPU: the oracle restatement against the fixtures produced by the reference itself."""

import pytest
import torch

from conftest import STEP_CASES, assert_close, assert_equal_exact
from humanoid_b200 import synth
from oracle import phc_oracle as O

# same torch build => bit-exact; a tiny tolerance keeps a different host ISA from breaking the pin
TIGHT = dict(rtol=1e-6, atol=1e-7)

MOTION_KEYS = (
    "root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa",
    "rg_pos", "rb_rot", "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights",
)  # fmt: skip


def _lib(g):
    return O.OracleMotionLib(g.group("in.lib"))


def test_frame_blend_bit_exact(golden):
    g = golden("frame_blend")
    i0, i1, bl = O.frame_blend(g.inp("time"), g.inp("len"), g.inp("num_frames"), g.inp("dt"))
    assert_equal_exact(i0, g.out("frame_idx0"), "frame_idx0")
    assert_equal_exact(i1, g.out("frame_idx1"), "frame_idx1")
    assert torch.equal(bl, g.out("blend"))


def test_slerp_and_exp_map(golden):
    g = golden("slerp")
    q = O.slerp(g.inp("q0"), g.inp("q1"), g.inp("t"))
    assert_close(q, g.out("q"), what="slerp", **TIGHT)
    assert_close(O.exp_map(g.inp("qn")), g.out("exp_map"), what="exp_map", **TIGHT)


@pytest.mark.parametrize("case", STEP_CASES)
def test_step_against_reference_fixture(golden, case):
    g = golden(case)
    lib = _lib(g)
    c = g.group("in.clock")
    T = int(g.inp("time_steps"))
    progress = c["progress_buf"].clone()
    obs, reward, raw, reset, term = O.step(
        lib, g.inp("state"), progress, c["motion_start_times"], c["motion_start_times_offset"],
        c["global_offset"], c["sampled_motion_ids"], g.inp("term_dist"), synth.SIM_DT,
        reset_body_ids=g.inp("reset_body_ids"), use_mean=bool(g.inp("use_mean")),
        enable_early_termination=bool(g.inp("early")), time_steps=T,
    )  # fmt: skip
    assert_equal_exact(progress, g.out("progress_after"), "progress")
    assert_equal_exact(reset, g.out("reset"), "reset")
    assert_equal_exact(term, g.out("terminated"), "terminated")
    assert_close(obs, g.out("obs"), what="obs", **TIGHT)
    assert_close(reward, g.out("reward"), what="reward", **TIGHT)
    assert_close(raw, g.out("reward_raw"), what="reward_raw", **TIGHT)
    # the fixture must actually exercise both outcomes somewhere in the suite
    assert obs.shape[1] == 358 + 576 * T


@pytest.mark.parametrize("case", STEP_CASES)
@pytest.mark.parametrize("which", ["t0", "t1"])
def test_motion_state_against_reference_fixture(golden, case, which):
    g = golden(case)
    lib = _lib(g)
    c = g.group("in.clock")
    res = lib.get_motion_state(c["sampled_motion_ids"], g.out(which), c["global_offset"])
    assert_equal_exact(res["frame_idx0"], g.out(f"{which}.frame_idx0"), "frame_idx0")
    assert_equal_exact(res["frame_idx1"], g.out(f"{which}.frame_idx1"), "frame_idx1")
    for k in MOTION_KEYS:
        assert_close(res[k], g.out(f"{which}.{k}"), what=k, **TIGHT)


def test_flag_variants(golden):
    g = golden("flags")
    pos, rot, vel, ang = synth.body_views(g.inp("state"))
    for name in ("default", "not_upright", "global_root", "no_height", "with_params", "all_off"):
        fl = [bool(x) for x in g.inp(f"self.{name}")]
        o = O.self_obs_smpl_max(pos, rot, vel, ang, g.inp("smpl"), g.inp("limb"), *fl)
        assert_close(o, g.out(f"self.{name}"), what=f"self obs {name}", **TIGHT)
    ref = g.group("in.ref")
    o = O.imitation_obs_v6(pos[:, 0], rot[:, 0], pos, rot, vel, ang,
                           ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], 1, False)  # fmt: skip
    assert_close(o, g.out("v6.not_upright"), what="v6 not upright", **TIGHT)
    s = g.inp("subset")
    o = O.imitation_obs_v6(pos[:, 0], rot[:, 0], pos[:, s], rot[:, s], vel[:, s], ang[:, s], ref["rg_pos"][:, s],
                           ref["rb_rot"][:, s], ref["body_vel"][:, s], ref["body_ang_vel"][:, s], 1, True)  # fmt: skip
    assert_close(o, g.out("v6.subset12"), what="v6 subset", **TIGHT)
    r, raw = O.imitation_reward(pos[:, 0], rot[:, 0], pos[:, s], rot[:, s], vel[:, s], ang[:, s], ref["rg_pos"][:, s],
                                ref["rb_rot"][:, s], ref["body_vel"][:, s], ref["body_ang_vel"][:, s],
                                O.DEFAULT_RWD_SPECS)  # fmt: skip
    assert_close(r, g.out("reward.subset12"), what="reward subset", **TIGHT)
    assert_close(raw, g.out("reward_raw.subset12"), what="reward_raw subset", **TIGHT)
    lib = _lib(g)
    res = lib.get_motion_state(g.inp("clock.sampled_motion_ids"), g.inp("t"), None)
    assert_close(res["rg_pos"], g.out("rg_pos.no_offset"), what="no offset", **TIGHT)


def test_v7_is_the_documented_v6_column_subset(golden):
    g = golden("step_T10")
    T = int(g.inp("time_steps"))
    cols = O.v7_columns(24, T)
    assert cols.numel() == T * 24 * 9
    v6 = g.out("task_obs")
    blk = v6.view(v6.shape[0], T, 576)
    want = torch.cat([blk[..., 0:72], blk[..., 216:288], blk[..., 360:432]], dim=-1).reshape(v6.shape[0], -1)
    assert torch.equal(v6[:, cols], want)


def test_running_norm(golden):
    g = golden("running_norm")
    m, v, c = torch.zeros(1, 934), torch.ones(1, 934), torch.ones(1)
    m, v, c = O.running_norm_update(m, v, c, g.inp("x1"))
    assert_close(m, g.out("mean1"), what="mean1", **TIGHT)
    assert_close(v, g.out("var1"), what="var1", **TIGHT)
    m, v, c = O.running_norm_update(m, v, c, g.inp("x2"))
    assert_close(m, g.out("mean2"), what="mean2", **TIGHT)
    assert_close(v, g.out("var2"), what="var2", **TIGHT)
    assert float(c) == float(g.out("count2"))
    assert_close(O.running_norm_forward(m, v, g.inp("x2")[:32]), g.out("fwd"), what="fwd", **TIGHT)


def test_episode_bookkeeping(golden):
    """oracle.episode_update against the recorded run of the reference's PHCPufferEnv.step."""
    from conftest import replay_episode_golden

    def new_state(n, cols):
        return dict(
            terminals=torch.zeros(n, dtype=torch.bool), truncations=torch.zeros(n, dtype=torch.bool),
            masks=torch.ones(n, dtype=torch.bool), episode_returns=torch.zeros(n),
            episode_lengths=torch.zeros(n, dtype=torch.int32), raw_rewards=torch.zeros(cols),
            stats=torch.zeros(4, dtype=torch.float64),
        )  # fmt: skip

    replay_episode_golden(golden("episode"), new_state, O.episode_update)


def test_env_rollout_vs_reference_step_and_reset(golden):
    """oracle.OracleEnv against a recording of the reference's own HumanoidPHC.step / reset methods."""
    from conftest import replay_env_rollout
    from util_cpu import oracle_env_from_golden

    g = golden("env_rollout")
    env = oracle_env_from_golden(g)
    assert replay_env_rollout(g, env) == 5
    assert sum(int(g.out(f"step.{k}.terminate").sum()) for k in range(5)) > 0


@pytest.mark.parametrize("mode", ["Default", "Hybrid"])
def test_env_reset_modes_vs_reference(golden, mode):
    """StateInit.Default / StateInit.Hybrid (envs/humanoid_phc.py:678-745): oracle.OracleEnv against a recording of the
    reference's own HumanoidPHC.step / reset in those modes."""
    from conftest import replay_env_reset_modes
    from util_cpu import oracle_env_from_golden

    g = golden("env_reset_modes")
    env = oracle_env_from_golden(g, use_amp_obs=False)
    env.initial_root_states, env.initial_dof_pos = g.inp("initial_root_states"), g.inp("initial_dof_pos")
    env.initial_dof_vel = g.inp("initial_dof_vel")
    assert replay_env_reset_modes(g, env, mode) == 3


def test_pd_action_offset_scale_vs_reference(golden):
    """The shim's host-side ``build_pd_action_offset_scale`` against the reference's own ``_build_pd_action_offset_scale``
    (envs/humanoid_phc.py:385-457) on random joint limits, all config variants.  1e-6: the reference multiplies numpy
    float32 scalars by Python floats, which NumPy 1 did in float64 and NumPy 2 does in float32."""
    from humanoid_b200.env import build_pd_action_offset_scale

    g = golden("pd_offset_scale")
    for name in ("default", "bias_offset", "smpl_pd_offset_upright", "smpl_pd_offset_not_upright"):
        bias, smpl_off, upright = (bool(x) for x in g.inp(name))
        off, sc = build_pd_action_offset_scale(g.inp("dof_limits_lower"), g.inp("dof_limits_upper"), bias, smpl_off, upright)
        assert_close(off, g.out(f"{name}.offset"), what=f"{name} offset", rtol=1e-6, atol=1e-6)
        assert_close(sc, g.out(f"{name}.scale"), what=f"{name} scale", rtol=1e-6, atol=1e-6)
        assert float(sc[1 * 3 + 1]) == 5.0 and float(sc[5 * 3 + 1]) == 5.0  # L_Knee, R_Knee: dofs 1 and 5, y axis


def test_fixtures_exercise_both_flag_values(golden):
    seen_reset, seen_term, seen_pass = set(), set(), set()
    for case in STEP_CASES:
        g = golden(case)
        seen_reset |= set(g.out("reset").tolist())
        seen_term |= set(g.out("terminated").tolist())
        seen_pass |= set(g.out("pass_time").tolist())
    assert seen_reset == {True, False} and seen_term == {True, False} and seen_pass == {True, False}


def test_sample_time_interval_bit_exact(golden):
    g = golden("sample_time")
    t = O.sample_time_interval(g.inp("motion_lengths"), g.inp("ids"), g.inp("phase"))
    assert torch.equal(t, g.out("motion_time"))
    k = t * 30
    assert float((k - k.round()).abs().max()) < 1e-3  # multiples of 1/30 s


def test_amp_obs_variants(golden):
    g = golden("amp_obs")
    a = {k: g.inp(k) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_pos", "dof_vel",
                                "key_body_pos", "shape", "limb", "dof_subset")}  # fmt: skip
    for name in ("default", "all_dofs", "global_root_no_height", "not_upright_with_params"):
        fl = [bool(x) for x in g.inp(f"flags.{name}")]
        o = O.amp_obs_smpl(a["root_pos"], a["root_rot"], a["root_vel"], a["root_ang_vel"], a["dof_pos"], a["dof_vel"],
                           a["key_body_pos"], a["shape"], a["limb"], a["dof_subset"], *fl)  # fmt: skip
        assert_close(o, g.out(name), what=f"amp obs {name}", **TIGHT)
    assert g.out("default").shape[1] == 196


# ---------------------------------------------------------------------------------------
# motion-library build (the load side): oracle/build_oracle.py against the reference's own load_motions
# ---------------------------------------------------------------------------------------
BUILD_KEYS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "grvs", "gravs", "motion_aa", "motion_lengths",
              "motion_num_frames", "motion_dt", "motion_fps", "length_starts", "motion_bodies", "motion_limb_weights")  # fmt: skip


def oracle_build_from_golden(g, variant):
    from oracle import build_oracle as B

    crop = variant == "cropped"  # load_motion_with_skeleton's max_length window (:773-778) at the recorded random starts
    return B.build_motion_library(
        g.inp("pose_quat_global").numpy(), g.inp("root_trans_offset").numpy(), g.inp("pose_aa").numpy(),
        g.inp("num_frames").tolist(), g.inp("fps").tolist(), g.inp("parent_indices").tolist(),
        g.inp("local_translation").numpy(), g.inp("gender_betas").numpy(), g.inp("limb_weights").numpy(),
        heading_u=None if variant == "deterministic" else g.inp("heading_u").tolist(),
        max_length=int(g.inp("crop_max_length")) if crop else -1, crop_start=g.inp("crop_start").tolist() if crop else None,
    )  # fmt: skip


@pytest.mark.parametrize("variant", ["deterministic", "random_heading", "cropped"])
def test_motion_build_against_reference_load_motions(golden, variant):
    g = golden("motion_build")
    torch.set_num_threads(1)
    out = oracle_build_from_golden(g, variant)
    for k in BUILD_KEYS:
        want = g.out(f"{variant}.{k}")
        assert out[k].dtype == want.dtype and out[k].shape == want.shape, k
        if want.dtype == torch.int64:
            assert_equal_exact(out[k], want, k)
        else:
            assert_close(out[k], want, what=f"{variant}.{k}", **TIGHT)
    # the fixture covers: a 2-frame clip, a 3-frame clip (shorter than the filter), a 60 fps clip, and crops that
    # start inside a clip
    assert g.inp("num_frames").min() == 2 and set(g.inp("fps").tolist()) == {30, 60}
    assert int(g.inp("crop_start").max()) > 0 and int(g.out("cropped.motion_num_frames").max()) == int(g.inp("crop_max_length"))
    # reference quirk: _motion_aa keeps the files' full length (motion_lib.py:377), so it is longer than gts when cropped
    assert g.out("cropped.motion_aa").shape[0] == int(g.inp("num_frames").sum()) > g.out("cropped.gts").shape[0]


def test_motion_build_host_constants_match_scipy(golden):
    """The two host-side constants the CUDA build takes: scipy's gaussian taps and the heading quaternion."""
    from scipy.ndimage import _filters
    from scipy.spatial.transform import Rotation as sRot

    from humanoid_b200 import motion_build as MB

    import numpy as np

    assert np.array_equal(MB.gaussian_taps(), _filters._gaussian_kernel1d(2.0, 0, 8))
    u = golden("motion_build").inp("heading_u").numpy()
    zw = MB.heading_half_angle(u)
    for i, ui in enumerate(u):
        q = sRot.from_euler("xyz", [0.0, 0.0, np.pi * (2 * ui - 1.0)]).as_quat()
        assert q[0] == 0 and q[1] == 0
        assert np.allclose(zw[i], q[2:], rtol=0, atol=2e-16)
    # the root rotation vector under the heading (motion_lib.py:794), incl. both small-angle branches
    rng = np.random.default_rng(0)
    v = rng.normal(size=(500, 3))
    v[::5] *= 1e-4
    v[::7] *= 3
    for ui in (0.1, 0.5, 0.93, 1e-4, 0.5 + 1e-5):
        h = sRot.from_euler("xyz", [0.0, 0.0, np.pi * (2 * ui - 1.0)])
        want = (h * sRot.from_rotvec(v)).as_rotvec()
        assert np.abs(MB.heading_on_rotvec(v, *MB.heading_half_angle(ui)) - want).max() < 4e-15
