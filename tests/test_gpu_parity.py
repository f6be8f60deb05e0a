"""GPU: the CUDA path (through the C ABI) against the reference-generated fixtures and the oracle."""

import pytest
import torch

from conftest import STEP_CASES, assert_close, assert_equal_exact

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from humanoid_b200 import (
        MotionLib,
        compute_humanoid_im_reset,
        compute_humanoid_observations_smpl_max,
        compute_imitation_observations_v6,
        compute_imitation_observations_v7,
        compute_imitation_reward,
        synth,
    )
    from oracle import phc_oracle as O
    from util_gpu import (DEV, DOF_DERIVED_TOL, DOF_TOL, OBS_TOL, clock_from_golden, env_from, lib_from_golden, make_case_cpu,
                          oracle_step)  # fmt: skip

MOTION_KEYS = (
    "root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa",
    "rg_pos", "rb_rot", "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights",
)  # fmt: skip


def cuda(x):
    return x.to(DEV)


# ---------------------------------------------------------------------------------------
# integer path
# ---------------------------------------------------------------------------------------
def test_frame_blend_bit_exact(golden):
    g = golden("frame_blend")
    lib = MotionLib(synth.make_motion_lib(2, 4, 6), device=DEV)
    i0, i1, bl = lib._calc_frame_blend(cuda(g.inp("time")), cuda(g.inp("len")), cuda(g.inp("num_frames")), cuda(g.inp("dt")))
    assert_equal_exact(i0, g.out("frame_idx0"), "frame_idx0")
    assert_equal_exact(i1, g.out("frame_idx1"), "frame_idx1")
    assert torch.equal(bl.cpu(), g.out("blend")), "blend is IEEE +-*/ only: must be bit-exact"


def test_frame_blend_env_clock_regime_bit_exact():
    """progress*dt + k/30 lands on frame boundaries (SURVEY §7 hard part 1): 1M random draws."""
    gen = torch.Generator().manual_seed(77)
    n = 1 << 20
    nf = torch.randint(2, 7000, (n,), generator=gen)
    fps = torch.tensor([30.0, 60.0, 120.0], dtype=torch.float64)[torch.randint(0, 3, (n,), generator=gen)]
    dt = (1.0 / fps).float()
    ln = ((1.0 / fps) * (nf - 1)).float()
    k = (torch.rand(n, generator=gen) * nf).long()
    p = torch.randint(0, 400, (n,), generator=gen).to(torch.int16)
    t = p * synth.SIM_DT + (k * (1 / 30)).float()
    want = O.frame_blend(t, ln, nf, dt)
    lib = MotionLib(synth.make_motion_lib(2, 4, 6), device=DEV)
    got = lib._calc_frame_blend(cuda(t), cuda(ln), cuda(nf), cuda(dt))
    assert_equal_exact(got[0], want[0], "frame_idx0")
    assert_equal_exact(got[1], want[1], "frame_idx1")
    assert torch.equal(got[2].cpu(), want[2])


# ---------------------------------------------------------------------------------------
# motion state
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", STEP_CASES)
@pytest.mark.parametrize("which", ["t0", "t1"])
def test_motion_state_vs_reference_fixture(golden, case, which):
    g = golden(case)
    lib = MotionLib(lib_from_golden(g), device=DEV)
    c = clock_from_golden(g)
    res = lib.get_motion_state(cuda(c.sampled_motion_ids), cuda(g.out(which)), cuda(c.global_offset), with_frame_info=True)
    assert_equal_exact(res["frame_idx0"], g.out(f"{which}.frame_idx0"), "frame_idx0")
    assert_equal_exact(res["frame_idx1"], g.out(f"{which}.frame_idx1"), "frame_idx1")
    assert torch.equal(res["blend"].cpu(), g.out(f"{which}.blend"))
    for k in MOTION_KEYS:
        tol = DOF_TOL if k == "dof_pos" else OBS_TOL
        assert_close(res[k], g.out(f"{which}.{k}"), what=f"{case}/{which}/{k}", **tol)
    for k in ("motion_aa", "motion_bodies", "motion_limb_weights"):  # pure gathers
        assert torch.equal(res[k].cpu(), g.out(f"{which}.{k}")), k


def test_motion_state_no_offset_and_negative_times(golden):
    g = golden("flags")
    lib_data = lib_from_golden(g)
    lib = MotionLib(lib_data, device=DEV)
    ids, t = g.inp("clock.sampled_motion_ids"), g.inp("t")
    res = lib.get_motion_state(cuda(ids), cuda(t), None)
    assert_close(res["rg_pos"], g.out("rg_pos.no_offset"), what="no offset", **OBS_TOL)
    t2 = t - 0.3
    want = O.OracleMotionLib(lib_data).get_motion_state(ids, t2, None)
    got = lib.get_motion_state(cuda(ids), cuda(t2), None, with_frame_info=True)
    assert_equal_exact(got["frame_idx0"], want["frame_idx0"], "idx0 (t<0)")
    assert_close(got["rb_rot"], want["rb_rot"], what="rb_rot (t<0)", **OBS_TOL)


def test_slerp_branches_through_motion_state(golden):
    """dot<0, |sin|<1e-3, |cos|>=1 (incl. non-unit inputs): one 2-frame clip per golden pair."""
    g = golden("slerp")
    q0, q1, t = g.inp("q0"), g.inp("q1"), g.inp("t").flatten()
    n = q0.shape[0]
    grs = torch.stack([q0, q1], dim=1).reshape(2 * n, 1, 4).expand(2 * n, 24, 4).contiguous()
    dt = 1.0 / 30.0
    data = synth.MotionData(
        gts=torch.zeros(2 * n, 24, 3), grs=grs, lrs=grs.clone(), gvs=torch.zeros(2 * n, 24, 3),
        gavs=torch.zeros(2 * n, 24, 3), dvs=torch.zeros(2 * n, 23, 3), motion_aa=torch.zeros(2 * n, 72),
        motion_lengths=torch.full((n,), dt), motion_num_frames=torch.full((n,), 2, dtype=torch.int64),
        motion_dt=torch.full((n,), dt), motion_fps=torch.full((n,), 30.0),
        length_starts=torch.arange(n, dtype=torch.int64) * 2, motion_bodies=torch.zeros(n, 17),
        motion_limb_weights=torch.zeros(n, 10),
    )  # fmt: skip
    lib = MotionLib(data, device=DEV)
    ids = torch.arange(n)
    times = (t * dt).float() * 0.999  # stay inside the clip so idx0 == 0
    res = lib.get_motion_state(cuda(ids), cuda(times), None, with_frame_info=True)
    assert int(res["frame_idx0"].max()) == 0
    blend = res["blend"].cpu()
    want = O.slerp(q0, q1, blend[:, None])
    assert not torch.isnan(want).any()
    assert_close(res["rb_rot"][:, 0], want, what="slerp", **OBS_TOL)
    assert_close(res["rb_rot"][:, 23], want, what="slerp (last body)", **OBS_TOL)
    want_dof = O.exp_map(O.slerp(q0, q1, blend[:, None]))
    assert_close(res["dof_pos"][:, 0:3], want_dof, what="exp_map", rtol=2e-4, atol=5e-5)
    # rows whose result is far from the identity are well conditioned: tight tolerance there
    far = want[:, 3].abs() < 0.99
    assert far.sum() > 100
    assert_close(res["dof_pos"][:, 0:3][far], want_dof[far], what="exp_map (well conditioned)", **OBS_TOL)


# ---------------------------------------------------------------------------------------
# whole step: fused kernel and the per-function decomposition, vs the reference fixtures
# ---------------------------------------------------------------------------------------
def _run_env(g, fused):
    T = int(g.inp("time_steps"))
    env = env_from(lib_from_golden(g), clock_from_golden(g), g.inp("state"), time_steps=T,
                   enable_early_termination=bool(g.inp("early")))  # fmt: skip
    env.set_termination_distances(cuda(g.inp("term_dist")))
    ids = g.inp("reset_body_ids")
    if ids.numel() != 24:
        env._reset_bodies_id = cuda(ids)
        env._step_args = None
    env.flag_im_eval = bool(g.inp("use_mean"))
    if fused:
        env.step()
    else:
        env.post_physics_step_unfused()
    torch.cuda.synchronize()
    return env


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "per_function"])
@pytest.mark.parametrize("case", STEP_CASES)
def test_step_vs_reference_fixture(golden, case, fused):
    g = golden(case)
    env = _run_env(g, fused)
    assert_equal_exact(env.progress_buf, g.out("progress_after"), "progress_buf")
    assert_equal_exact(env.reset_buf, g.out("reset"), "reset_buf")
    assert_equal_exact(env._terminate_buf, g.out("terminated"), "_terminate_buf")
    assert_close(env.obs_buf, g.out("obs"), what="obs_buf", **OBS_TOL)
    assert_close(env.rew_buf, g.out("reward"), what="rew_buf", **OBS_TOL)
    assert_close(env.reward_raw[:, :4], g.out("reward_raw"), what="reward_raw", **OBS_TOL)


def test_flag_variants_vs_reference_fixture(golden):
    g = golden("flags")
    state = cuda(g.inp("state"))
    pos, rot, vel, ang = synth.body_views(state)
    smpl, limb = cuda(g.inp("smpl")), cuda(g.inp("limb"))
    for name in ("default", "not_upright", "global_root", "no_height", "with_params", "all_off"):
        fl = [bool(x) for x in g.inp(f"self.{name}")]
        o = compute_humanoid_observations_smpl_max(pos, rot, vel, ang, smpl, limb, *fl)
        assert_close(o, g.out(f"self.{name}"), what=f"self obs {name}", **OBS_TOL)
    ref = {k: cuda(v) for k, v in g.group("in.ref").items()}
    o = compute_imitation_observations_v6(pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"],
                                          ref["body_vel"], ref["body_ang_vel"], 1, False)  # fmt: skip
    assert_close(o, g.out("v6.not_upright"), what="v6 not upright", **OBS_TOL)
    s = cuda(g.inp("subset"))
    args = (pos[:, 0], rot[:, 0], pos[:, s], rot[:, s], vel[:, s], ang[:, s], ref["rg_pos"][:, s], ref["rb_rot"][:, s],
            ref["body_vel"][:, s], ref["body_ang_vel"][:, s])  # fmt: skip
    assert_close(compute_imitation_observations_v6(*args, 1, True), g.out("v6.subset12"), what="v6 subset", **OBS_TOL)
    r, raw = compute_imitation_reward(*args, O.DEFAULT_RWD_SPECS)
    assert_close(r, g.out("reward.subset12"), what="reward subset", **OBS_TOL)
    assert_close(raw, g.out("reward_raw.subset12"), what="reward_raw subset", **OBS_TOL)


def test_v7_is_the_v6_column_subset(golden):
    g = golden("step_T10")
    T = int(g.inp("time_steps"))
    state = cuda(g.inp("state"))
    pos, rot, vel, ang = synth.body_views(state)
    lib_data, c = lib_from_golden(g), clock_from_golden(g)
    ol = O.OracleMotionLib(lib_data)
    p = c.progress_buf + 1
    refs = [ol.get_motion_state(c.sampled_motion_ids, (p + k) * synth.SIM_DT + c.motion_start_times
                                + c.motion_start_times_offset, c.global_offset) for k in range(1, T + 1)]  # fmt: skip
    st = lambda key: cuda(torch.stack([r[key] for r in refs], 1).flatten(0, 1))  # noqa: E731
    args = (pos[:, 0], rot[:, 0], pos, rot, vel, ang, st("rg_pos"), st("rb_rot"), st("body_vel"), st("body_ang_vel"), T, True)
    v6 = compute_imitation_observations_v6(*args)
    v7 = compute_imitation_observations_v7(*args)
    assert_close(v6, g.out("task_obs"), what="v6 T=10", **OBS_TOL)
    assert torch.equal(v7, v6[:, cuda(O.v7_columns(24, T))])


# ---------------------------------------------------------------------------------------
# larger seeded workloads vs the oracle (BASELINE configs 1, 2, 4, 5 at oracle-friendly sizes)
# ---------------------------------------------------------------------------------------
CASES = {
    "config1_N256_M16": dict(num_envs=256, num_motions=16, seed=101, max_progress=30),
    "config2_N4096_ids_arange": dict(num_envs=4096, num_motions=4096, seed=102, max_frames=120, max_progress=40),
    "config4_random_ids_mixed_fps": dict(num_envs=2048, num_motions=96, seed=103, ids="random", aligned=False,
                                         fps_choices=(30, 60, 120), min_frames=60, max_frames=900, max_progress=40),
    "random_rotations": dict(num_envs=1024, num_motions=64, seed=104, rot_regime="random", max_progress=40),
    "ragged_N1001": dict(num_envs=1001, num_motions=37, seed=105, max_progress=40),
    "tiny_N1": dict(num_envs=1, num_motions=1, seed=106, max_progress=5),
}  # fmt: skip


@pytest.mark.parametrize("name", list(CASES))
def test_fused_step_vs_oracle(name):
    lib_data, clock, state = make_case_cpu(**CASES[name])
    obs, rew, raw, reset, term, prog = oracle_step(lib_data, clock, state)
    env = env_from(lib_data, clock, state)
    env.step()
    assert_equal_exact(env.progress_buf, prog, "progress")
    assert_equal_exact(env.reset_buf, reset, f"{name}: reset")
    assert_equal_exact(env._terminate_buf, term, f"{name}: terminated")
    assert_close(env.obs_buf, obs, what=f"{name}: obs", **OBS_TOL)
    assert_close(env.rew_buf, rew, what=f"{name}: reward", **OBS_TOL)
    assert_close(env.reward_raw[:, :4], raw, what=f"{name}: reward_raw", **OBS_TOL)
    if state.shape[0] >= 256:
        assert 0.02 < term.float().mean() < 0.98, "workload should mix terminated / alive envs"


def test_fused_step_T10_vs_oracle():
    lib_data, clock, state = make_case_cpu(num_envs=512, num_motions=64, seed=107, max_progress=40)
    obs, rew, raw, reset, term, prog = oracle_step(lib_data, clock, state, time_steps=10)
    env = env_from(lib_data, clock, state, time_steps=10)
    env.step()
    assert env.obs_buf.shape == (512, 358 + 5760)
    assert_close(env.obs_buf, obs, what="obs T=10", **OBS_TOL)
    assert_equal_exact(env.reset_buf, reset, "reset")
    assert_close(env.rew_buf, rew, what="reward", **OBS_TOL)


def test_eval_mode_reset_vs_oracle():
    lib_data, clock, state = make_case_cpu(num_envs=2048, num_motions=32, seed=108, max_progress=40)
    eval_ids = torch.tensor([i for i in range(24) if i not in (4, 8, 18, 23)])
    o = oracle_step(lib_data, clock, state, term=0.12, reset_body_ids=eval_ids, use_mean=True)
    env = env_from(lib_data, clock, state)
    env.toggle_eval_mode()
    env.set_termination_distances(0.12)
    env.step()
    # the mean over the 20 eval bodies is summed in ATen's association (phc_math.cuh aten_row_sum): flags are exact
    assert_equal_exact(env._terminate_buf, o[4], "eval-mode terminated")
    assert_equal_exact(env.reset_buf, o[3], "eval-mode reset")
    assert 0.05 < o[4].float().mean() < 0.95
    # extras["mpjpe"] (humanoid_phc.py:159-167): mean over all 24 bodies at the reward time
    t = synth.reward_time(clock, extra_steps=1)
    ref = O.OracleMotionLib(lib_data).get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    want = (state[:, :, 0:3] - ref["rg_pos"]).norm(dim=-1).mean(dim=-1)
    assert_close(env.extras["mpjpe"], want, what="mpjpe", **OBS_TOL)


def test_eval_mode_termination_on_threshold_rows_bit_exact(golden):
    """compute_humanoid_im_reset(use_mean=True) through K5 on rows whose mean distance sits within a few ulps of the
    threshold, for every reset-body count: the fixture is the reference's own output, every flag must be equal."""
    g = golden("reset_mean")
    thr = float(g.inp("threshold"))
    for R in g.inp("rs").tolist():
        pos, ref = cuda(g.inp(f"pos{R}")), cuda(g.inp(f"ref{R}"))
        n = pos.shape[0]
        got = compute_humanoid_im_reset(
            torch.ones(n, dtype=torch.bool, device=DEV), cuda(g.inp(f"progress{R}")), None, None, pos, ref,
            cuda(g.inp(f"pass_time{R}")), True, torch.full((R,), thr, device=DEV), True)  # fmt: skip
        assert_equal_exact(got[1], g.out(f"terminated{R}"), f"terminated, R = {R}")
        assert_equal_exact(got[0], g.out(f"reset{R}"), f"reset, R = {R}")


@pytest.mark.parametrize("T,generic", [(1, False), (1, True), (3, False)], ids=["fast", "generic", "multi"])
def test_eval_mode_fused_step_flags_on_threshold_rows_bit_exact(T, generic):
    """The fused kernels' eval-mode test on envs posed so that the mean distance of the 20 eval bodies hugs the
    threshold (to within a few floats either side, in the oracle's arithmetic): flags equal to the oracle's."""
    from humanoid_b200 import _cabi

    N, thr = 2048, 0.2
    lib_data, clock, state = make_case_cpu(num_envs=N, num_motions=64, seed=131, max_progress=40)
    eval_ids = torch.tensor([i for i in range(24) if i not in (4, 8, 18, 23)])
    t = synth.reward_time(clock, extra_steps=1)
    ref = O.OracleMotionLib(lib_data).get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)["rg_pos"]
    gen = torch.Generator().manual_seed(7)
    delta = torch.randn(N, 24, 3, generator=gen) * 0.1
    for _ in range(4):
        m = torch.norm(((ref + delta) - ref)[:, eval_ids], dim=-1).mean(dim=-1)
        delta = delta * (thr / m)[:, None, None]
    delta = delta * (1.0 + torch.randint(-2, 4, (N,), generator=gen).float() * 6e-8)[:, None, None]
    state = state.clone()
    state[:, :, 0:3] = ref + delta
    o = oracle_step(lib_data, clock, state, term=thr, reset_body_ids=eval_ids, use_mean=True, time_steps=T)
    mean = torch.norm((state[:, :, 0:3] - ref)[:, eval_ids], dim=-1).mean(dim=-1)
    assert ((mean - thr).abs() < 1e-7).float().mean() > 0.8, "rows should sit on the threshold"
    assert 0.2 < o[4].float().mean() < 0.8
    env = env_from(lib_data, clock, state, time_steps=T)
    env.toggle_eval_mode()
    env.set_termination_distances(thr)
    _cabi.load().phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if generic else 0)
    try:
        env.step()
    finally:
        _cabi.load().phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    assert_equal_exact(env._terminate_buf, o[4], "terminated")
    assert_equal_exact(env.reset_buf, o[3], "reset")


# ---------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes (no oracle needed)
# ---------------------------------------------------------------------------------------
def _gpu_case(N, M, seed, **kw):
    def query(lib_data, ids, times, offset):
        return MotionLib(lib_data, device=DEV).get_motion_state(ids, times, offset)

    return synth.make_case(N, M, query, seed=seed, device=DEV, **kw)


@pytest.mark.parametrize("N,M", [(4096, 4096), (65536, 8192)])
def test_fused_equals_per_function_kernels_bitwise(N, M):
    lib_data, clock, state = _gpu_case(N, M, 201, max_frames=90, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    a = HumanoidPHC(lib, N, device=DEV)
    b = HumanoidPHC(lib, N, device=DEV)
    for env in (a, b):
        env.set_sim_state(state)
        env.set_clock(clock)
    a.step()
    b.post_physics_step_unfused()
    assert torch.equal(a.obs_buf, b.obs_buf)
    assert torch.equal(a.rew_buf, b.rew_buf)
    assert torch.equal(a.reward_raw[:, :4], b.reward_raw[:, :4])
    assert torch.equal(a.reset_buf, b.reset_buf) and torch.equal(a._terminate_buf, b._terminate_buf)
    assert torch.equal(a.progress_buf, b.progress_buf)
    assert 0.02 < a._terminate_buf.float().mean() < 0.98


def test_env_partition_invariance_bitwise():
    """Multi-GPU contract: envs [lo,hi) computed alone == the same rows of the whole batch."""
    N = 16384
    lib_data, clock, state = _gpu_case(N, 2048, 202, max_frames=90, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    whole = HumanoidPHC(lib, N, device=DEV)
    whole.set_sim_state(state)
    whole.set_clock(clock)
    whole.step()
    for lo, hi in ((0, 8192), (8192, 16384), (4099, 9001)):
        part = HumanoidPHC(lib, hi - lo, device=DEV)
        part.set_sim_state(state[lo:hi])
        part.set_clock(synth.Clock(**{k: v[lo:hi] for k, v in clock.__dict__.items()}))
        part.step()
        assert torch.equal(part.obs_buf, whole.obs_buf[lo:hi])
        assert torch.equal(part.rew_buf, whole.rew_buf[lo:hi])
        assert torch.equal(part.reset_buf, whole.reset_buf[lo:hi])


def test_strided_views_and_contiguous_inputs_agree():
    lib_data, clock, state = _gpu_case(1536, 64, 203, max_progress=20)
    wide = torch.zeros(1536, 27, 13, device=DEV)  # extra actors per env, as bodies_per_env > num_bodies
    wide[:, :24] = state
    pos, rot, vel, ang = synth.body_views(wide)
    a = compute_humanoid_observations_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    b = compute_humanoid_observations_smpl_max(pos.contiguous(), rot.contiguous(), vel.contiguous(), ang.contiguous(),
                                               None, None, True, True, True, False, False)  # fmt: skip
    assert torch.equal(a, b)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    e1 = HumanoidPHC(lib, 1536, device=DEV)
    e1.set_sim_state(state)
    e1.set_clock(clock)
    e1.step()
    e2 = HumanoidPHC(lib, 1536, device=DEV, bodies_per_env=27)  # env stride 351: not 16-B aligned -> strided path
    e2.set_sim_state(wide)
    e2.set_clock(clock)
    e2.step()
    assert torch.equal(e1.obs_buf, e2.obs_buf) and torch.equal(e1.rew_buf, e2.rew_buf)


def test_subset_env_ids_observations():
    lib_data, clock, state = _gpu_case(512, 32, 204, max_progress=20)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    env = HumanoidPHC(lib, 512, device=DEV)
    env.set_sim_state(state)
    env.set_clock(clock)
    full = env._compute_observations().clone()
    env.obs_buf.zero_()
    ids = torch.tensor([3, 17, 200, 511], device=DEV)
    sub = env._compute_observations(ids)
    assert torch.equal(sub, full[ids]) and torch.equal(env.obs_buf[ids], full[ids])
    assert float(env.obs_buf[0].abs().sum()) == 0.0


def test_multi_step_trajectory_vs_oracle():
    """Five consecutive steps with the clock advancing in-kernel."""
    lib_data, clock, state = make_case_cpu(num_envs=300, num_motions=300, seed=109, min_frames=20, max_frames=40)
    env = env_from(lib_data, clock, state)
    ol = O.OracleMotionLib(lib_data)
    prog = clock.progress_buf.clone()
    for it in range(5):
        want = O.step(ol, state, prog, clock.motion_start_times, clock.motion_start_times_offset, clock.global_offset,
                      clock.sampled_motion_ids, torch.full((24,), 0.25), synth.SIM_DT)  # fmt: skip
        env.step()
        assert_equal_exact(env.progress_buf, prog, f"progress @{it}")
        assert_equal_exact(env.reset_buf, want[3], f"reset @{it}")
        assert_equal_exact(env._terminate_buf, want[4], f"terminated @{it}")
        assert_close(env.obs_buf, want[0], what=f"obs @{it}", **OBS_TOL)
        assert_close(env.rew_buf, want[1], what=f"reward @{it}", **OBS_TOL)


def test_step_is_cuda_graph_capturable():
    lib_data, clock, state = _gpu_case(2048, 128, 205, max_progress=10)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    ref_env = HumanoidPHC(lib, 2048, device=DEV)
    ref_env.set_sim_state(state)
    ref_env.set_clock(clock)
    for _ in range(3):
        ref_env.step()
    env = HumanoidPHC(lib, 2048, device=DEV)
    env.set_sim_state(state)
    env.set_clock(clock)
    env.step()  # warm up (sets the smem attribute) outside capture
    env.set_clock(clock)
    graph = torch.cuda.CUDAGraph()
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            env.step()
    torch.cuda.synchronize()
    env.set_clock(clock)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.progress_buf, ref_env.progress_buf)
    assert torch.equal(env.obs_buf, ref_env.obs_buf) and torch.equal(env.rew_buf, ref_env.rew_buf)


def test_standalone_reset_and_reward_vs_oracle():
    lib_data, clock, state = make_case_cpu(num_envs=777, num_motions=50, seed=110, max_progress=30)
    ol = O.OracleMotionLib(lib_data)
    t = synth.reward_time(clock, extra_steps=1)
    ref = ol.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    pos, rot, vel, ang = synth.body_views(state)
    prog = (clock.progress_buf + 1).to(torch.int16)
    pass_time = t >= lib_data.motion_lengths[clock.sampled_motion_ids]
    term = torch.full((24,), 0.25)
    want = O.im_reset(torch.ones(777, dtype=torch.bool), prog, None, None, pos.clone(), ref["rg_pos"], pass_time, True, term, False)
    sc = cuda(state)
    p, r, v, a = synth.body_views(sc)
    got = compute_humanoid_im_reset(torch.ones(777, dtype=torch.bool, device=DEV), cuda(prog), None, None, p,
                                    cuda(ref["rg_pos"]), cuda(pass_time), True, cuda(term), False)  # fmt: skip
    assert got[0].dtype == torch.bool
    assert_equal_exact(got[0], want[0], "reset")
    assert_equal_exact(got[1], want[1], "terminated")
    wr = O.imitation_reward(pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"], ref["body_vel"],
                            ref["body_ang_vel"], O.DEFAULT_RWD_SPECS)  # fmt: skip
    gr = compute_imitation_reward(p[:, 0], r[:, 0], p, r, v, a, cuda(ref["rg_pos"]), cuda(ref["rb_rot"]),
                                  cuda(ref["body_vel"]), cuda(ref["body_ang_vel"]), O.DEFAULT_RWD_SPECS)  # fmt: skip
    assert_close(gr[0], wr[0], what="reward", **OBS_TOL)
    assert_close(gr[1], wr[1], what="reward_raw", **OBS_TOL)


def test_empty_batch_is_a_no_op():
    lib = MotionLib(synth.make_motion_lib(3, 5, 9), device=DEV)
    res = lib.get_motion_state(torch.zeros(0, dtype=torch.int64, device=DEV), torch.zeros(0, device=DEV), None)
    assert res["rg_pos"].shape == (0, 24, 3)
    z = torch.zeros(0, 24, 13, device=DEV)
    o = compute_humanoid_observations_smpl_max(*synth.body_views(z), None, None, True, True, True, False, False)
    assert o.shape == (0, 358)


@pytest.mark.parametrize("N", [4096, 4099, 7])
def test_fast_kernel_equals_generic_kernel_bitwise(N):
    """T=1 dispatches to the TMA kernel (4 or 8 envs per block); the generic kernel is the same math."""
    from humanoid_b200 import HumanoidPHC, _cabi

    lib_data, clock, state = _gpu_case(N, max(N // 4, 3), 206, max_frames=90, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    outs = []
    try:
        for generic, epb in ((1, 4), (0, 4)):
            assert capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, generic) == 0
            assert capi.phc_set_option(_cabi.OPT_STEP_EPB, epb) == 0
            env = HumanoidPHC(lib, N, device=DEV, obs_moments=True)
            env.set_sim_state(state)
            env.set_clock(clock)
            env.step()
            torch.cuda.synchronize()
            outs.append(env)
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
        capi.phc_set_option(_cabi.OPT_STEP_EPB, 4)
    g = outs[0]
    for f in outs[1:]:
        assert torch.equal(f.obs_buf, g.obs_buf)
        assert torch.equal(f.rew_buf, g.rew_buf) and torch.equal(f.reward_raw, g.reward_raw)
        assert torch.equal(f.reset_buf, g.reset_buf) and torch.equal(f._terminate_buf, g._terminate_buf)
        assert torch.equal(f.progress_buf, g.progress_buf)
        torch.testing.assert_close(f.obs_moments, g.obs_moments, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("N,regime", [(8192, "aligned30"), (16387, "aligned30"), (6150, "mixed_fps_unaligned"), (7, "aligned30"),
                                      (2401, "power")])  # fmt: skip
def test_persistent_kernel_equals_generic_kernel_bitwise(N, regime):
    """K6-persist (persistent blocks, producer warp + consumer warps, tiles drawn from a device counter) against the
    generic kernel: every output bit for bit, three steps deep (the counter must rewind itself between launches),
    ragged N, the two-span frame case of unaligned clips (with three frame slots per env: the far row read from the
    packed table), and the power reward; both instantiations (3 slots / 5 blocks per SM, 4 slots / 4 blocks)."""
    from humanoid_b200 import HumanoidPHC, _cabi

    kw = dict(max_progress=40)
    if regime == "mixed_fps_unaligned":
        kw.update(ids="random", aligned=False, fps_choices=(30, 60, 120), min_frames=60, max_frames=400)
    lib_data, clock, state = _gpu_case(N, 300, 77, **kw)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    outs = []
    try:
        for which in ("persist_3slots_5blocks", "persist_4slots_4blocks", "generic"):
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, {"persist_3slots_5blocks": 2, "persist_4slots_4blocks": 3}.get(which, 0))
            capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if which == "generic" else 0)
            env = HumanoidPHC(lib, N, device=DEV, use_power_reward=regime == "power")
            env.set_sim_state(state)
            env.set_clock(clock)
            if regime == "power":
                env.dof_force_tensor.normal_(generator=torch.Generator(device=DEV).manual_seed(1))
                env._dof_vel.copy_(torch.randn(N, 69, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV))
            snaps = []
            for _ in range(3):
                env.step()
                snaps.append([getattr(env, k).clone() for k in ("obs_buf", "rew_buf", "reward_raw", "reset_buf",
                                                                "_terminate_buf", "progress_buf", "_reset_out", "_terminate_out")])
            torch.cuda.synchronize()
            outs.append(snaps)
    finally:
        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    for variant in (0, 1):
        for k, (a, b) in enumerate(zip(outs[variant], outs[2])):
            for x, y, nm in zip(a, b, ("obs", "rew", "reward_raw", "reset", "terminate", "progress", "reset_out", "terminate_out")):
                assert torch.equal(x, y), f"variant {variant}, step {k}: {nm}"


@pytest.mark.parametrize("variant", [2, 3], ids=["3slots_5blocks", "4slots_4blocks"])
def test_persistent_kernel_with_resets_between_steps_equals_single_wave_kernel(variant):
    """Four steps with a device-side reset of the flagged envs in between (their clocks restart under the producers'
    three-tile pipeline), mixed 30 / 60 fps clips and unaligned times (far rows): every output bit-identical to the
    single-wave kernel's."""
    from humanoid_b200 import HumanoidPHC, _cabi

    N = 16390
    lib_data, clock, state = _gpu_case(N, 300, 79, max_progress=30, ids="random", aligned=False, fps_choices=(30, 60),
                                       min_frames=80, max_frames=300)  # fmt: skip
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    outs = []
    gen = torch.Generator().manual_seed(3)
    phases = [torch.rand(N, generator=gen).to(DEV) for _ in range(4)]
    try:
        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, variant)
        for _ in range(2):
            env = HumanoidPHC(lib, N, device=DEV)
            env.set_sim_state(state)
            env.set_clock(clock)
            for k in range(4):
                env.step()
                if k == 1:
                    env.reset_done(phases[k])  # clocks of the flagged envs change under the next step's speculation
            torch.cuda.synchronize()
            outs.append(env)
    finally:
        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 0)
    ref = HumanoidPHC(lib, N, device=DEV)  # the single-wave kernel as the reference
    ref.set_sim_state(state)
    ref.set_clock(clock)
    for k in range(4):
        ref.step()
        if k == 1:
            ref.reset_done(phases[k])
    capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    for f in outs:
        for nm in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf", "progress_buf"):
            assert torch.equal(getattr(f, nm), getattr(ref, nm)), nm


def test_persistent_kernel_in_a_cuda_graph_with_many_launches():
    """300 consecutive launches in one graph (more than the 256 tile-counter slots), replayed twice, against the
    single-wave kernel: the counters rewind themselves and a slot that comes round again is clean."""
    from humanoid_b200 import HumanoidPHC, _cabi

    N = 6200
    lib_data, clock, state = _gpu_case(N, 200, 78, max_progress=10, min_frames=400, max_frames=500)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    res = []
    try:
        for mode in (2, 0):
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, mode)
            env = HumanoidPHC(lib, N, device=DEV)
            env.set_sim_state(state)
            env.set_clock(clock)
            env.step()
            env.set_clock(clock)
            torch.cuda.synchronize()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    for _ in range(300):
                        env.post_physics_step(True)
            for _ in range(2):
                env.set_clock(clock)
                torch.cuda.synchronize()
                g.replay()
            torch.cuda.synchronize()
            res.append([env.obs_buf.clone(), env.rew_buf.clone(), env.progress_buf.clone(), env.reset_buf.clone()])
    finally:
        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    for x, y in zip(*res):
        assert torch.equal(x, y)
    assert int(res[0][2][0]) == int(clock.progress_buf[0]) + 300


def test_fused_moments_epilogue_matches_column_moments():
    from humanoid_b200 import HumanoidPHC, RunningNorm

    lib_data, clock, state = _gpu_case(3000, 100, 207, max_progress=40)
    env = HumanoidPHC(MotionLib(lib_data, device=DEV), 3000, device=DEV, obs_moments=True)
    env.set_sim_state(state)
    env.set_clock(clock)
    env.step()
    rn = RunningNorm(934, device=DEV)
    sums = rn.moments(env.obs_buf)
    torch.testing.assert_close(env.obs_moments, sums, rtol=1e-12, atol=1e-9)
    x = env.obs_buf.double()
    torch.testing.assert_close(sums[:934], x.sum(0), rtol=1e-12, atol=1e-9)
    # the step's blocks spread over buckets (one accumulator would serialise the fp64 atomics); a rollout folds them
    assert env._obs_moment_buckets.shape == (32, 2 * 934) and int((env._obs_moment_buckets[:, 0] != 0).sum()) == 32
    env.step()
    both = torch.cat([x, env.obs_buf.double()])
    got, rows = env.take_obs_moments()
    assert rows == 2 * 3000 and float(env._obs_moment_buckets.abs().max()) == 0.0 and env.obs_moment_rows == 0
    torch.testing.assert_close(got[:934], both.sum(0), rtol=1e-12, atol=1e-9)
    torch.testing.assert_close(got[934:], (both * both).sum(0), rtol=1e-12, atol=1e-9)
    got2, _ = env.take_obs_moments(got)  # += into the caller's accumulator; nothing new since the last take
    assert got2 is got and torch.equal(got2[:934], got[:934])
    # the standalone kernel on a rollout-sized buffer (more rows per block, still one atomic per column per block)
    big = torch.randn(70001, 934, device=DEV) * 3 + 1
    s_big = rn.moments(big)
    torch.testing.assert_close(s_big[:934], big.double().sum(0), rtol=1e-12, atol=1e-8)
    torch.testing.assert_close(s_big[934:], (big.double() ** 2).sum(0), rtol=1e-12, atol=1e-7)
    torch.testing.assert_close(sums[934:], (x * x).sum(0), rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("bulk", [1, 0], ids=["bulk_reduce", "atomics"])
@pytest.mark.parametrize("N", [4096, 4099, 33, 5])
def test_moments_epilogue_bulk_reductions_equal_column_moments(N, bulk):
    """The T = 1 step's moments epilogue stages each block's 1868 fp64 partial sums in dead shared memory and hands
    them to the TMA engine as three bulk reductions (cp.reduce.async.bulk .add.f64) instead of issuing 1868 atomics.
    Both ways, ragged N included: partials equal phc_obs_moments to 1e-12 relative, every other output is untouched."""
    from humanoid_b200 import HumanoidPHC, RunningNorm, _cabi

    lib_data, clock, state = _gpu_case(N, 64, 41, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    plain = HumanoidPHC(lib, N, device=DEV)
    env = HumanoidPHC(lib, N, device=DEV, obs_moments=True)
    for e in (plain, env):
        e.set_sim_state(state)
        e.set_clock(clock)
    capi = _cabi.load()
    assert capi.phc_set_option(_cabi.OPT_MOMENTS_BULK, bulk) == 0
    try:
        for _ in range(2):
            plain.step()
            env.step()
        torch.cuda.synchronize()
    finally:
        capi.phc_set_option(_cabi.OPT_MOMENTS_BULK, 0)
    for k in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf", "progress_buf"):
        assert torch.equal(getattr(env, k), getattr(plain, k)), k
    rn = RunningNorm(934, device=DEV)
    # two steps were accumulated; the second step's rows are in obs_buf, the first step's are recomputed
    redo = HumanoidPHC(lib, N, device=DEV)
    redo.set_sim_state(state)
    redo.set_clock(clock)
    redo.step()
    want = rn.moments(redo.obs_buf)
    rn.moments(env.obs_buf, want)
    got, rows = env.take_obs_moments()
    assert rows == 2 * N
    scale = want.abs().clamp_min(1.0)
    assert float(((got - want).abs() / scale).max()) < 1e-12


@pytest.mark.parametrize("N,regime", [(8192, "aligned30"), (16387, "aligned30"), (9001, "mixed_fps_unaligned"), (5, "aligned30"),
                                      (2371, "aligned30")])  # fmt: skip
def test_persistent_kernel_keeps_the_moments_in_registers(N, regime):
    """With obs_moments the persistent kernel keeps each block's RunningNorm partials in registers across its tiles
    (ten fp64 column sums + ten sums of squares per consumer thread) and adds them to the accumulator buckets once at
    the end.  Every other output bit-identical to the plain step; partials equal phc_obs_moments of the rows to 1e-12
    relative; three steps deep (the partials of consecutive launches add up), ragged N, far rows."""
    from humanoid_b200 import HumanoidPHC, RunningNorm, _cabi

    kw = dict(max_progress=40)
    if regime == "mixed_fps_unaligned":
        kw.update(ids="random", aligned=False, fps_choices=(30, 60, 120), min_frames=60, max_frames=400)
    lib_data, clock, state = _gpu_case(N, 300, 91, **kw)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    rn = RunningNorm(934, device=DEV)
    plain = HumanoidPHC(lib, N, device=DEV)
    env = HumanoidPHC(lib, N, device=DEV, obs_moments=True)
    for e in (plain, env):
        e.set_sim_state(state)
        e.set_clock(clock)
    want = torch.zeros(2 * 934, dtype=torch.float64, device=DEV)
    try:
        for _ in range(3):
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 0)
            plain.step()
            capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 2)  # the persistent kernel whatever N
            env.step()
            rn.moments(plain.obs_buf, want)
            for k in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf", "progress_buf"):
                assert torch.equal(getattr(env, k), getattr(plain, k)), k
    finally:
        capi.phc_set_option(_cabi.OPT_STEP_PERSIST, 1)
    got, rows = env.take_obs_moments()
    assert rows == 3 * N
    scale = want.abs().clamp_min(1.0)
    assert float(((got - want).abs() / scale).max()) < 1e-12
    assert float(env._obs_moment_buckets.abs().max()) == 0.0


def test_running_norm_vs_reference_fixture(golden):
    from humanoid_b200 import RunningNorm

    g = golden("running_norm")
    rn = RunningNorm(934, device=DEV)
    rn.update(cuda(g.inp("x1")))
    assert_close(rn.running_mean, g.out("mean1"), what="mean1", rtol=1e-5, atol=1e-6)
    assert_close(rn.running_var, g.out("var1"), what="var1", rtol=1e-5, atol=1e-6)
    rn.update(cuda(g.inp("x2")))
    assert_close(rn.running_mean, g.out("mean2"), what="mean2", rtol=1e-5, atol=1e-6)
    assert_close(rn.running_var, g.out("var2"), what="var2", rtol=1e-5, atol=1e-6)
    assert float(rn.count) == float(g.out("count2"))
    assert_close(rn(cuda(g.inp("x2")[:32])), g.out("fwd"), what="forward", rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["pinned_direct", "pinned_staged", "pinned_auto", "pinned_tuned", "pageable_staged"])
def test_host_pipeline_matches_device_path(mode, monkeypatch):
    """phc_host_step on host buffers == the device path, bit for bit, on every schedule: pinned buffers with the kernels
    posting into the mapped host memory (direct), pinned buffers with copy-engine D2H (staged), the automatic choice
    between the two with a fixed chunk count (timed over the first calls), the full tuning of path and chunk count
    (num_chunks = 0: every candidate runs, every call is checked), and pageable buffers (always staged)."""
    import ctypes as C

    from humanoid_b200 import HumanoidPHC, _cabi

    pinned = mode != "pageable_staged"
    if mode in ("pinned_direct", "pinned_staged"):
        monkeypatch.setenv("PHC_HOST_PATH", mode.split("_")[1])
    else:
        monkeypatch.delenv("PHC_HOST_PATH", raising=False)
    N = 5000
    lib_data, clock, state = _gpu_case(N, 64, 208, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    env = HumanoidPHC(lib, N, device=DEV)
    env.set_sim_state(state)
    env.set_clock(clock)
    env.step()
    capi = _cabi.load()
    ctx = C.c_void_p()
    term = (C.c_float * 24)(*([0.25] * 24))
    spec = _cabi.reward_spec(env.rwd_specs)
    want_chunks = 0 if mode == "pinned_tuned" else 3
    _cabi.check(capi.phc_host_step_create(lib.handle, N, 1, want_chunks, term, 0xFFFFFF, 0, 1, synth.SIM_DT, C.byref(spec),
                                          C.byref(ctx)), "create")  # fmt: skip
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h = dict(
        state=mk(state.cpu().contiguous()), prog=mk(clock.progress_buf.cpu().clone()),
        start=mk(clock.motion_start_times.cpu()), off=mk(clock.motion_start_times_offset.cpu()),
        goff=mk(clock.global_offset.cpu().contiguous()), ids=mk(clock.sampled_motion_ids.cpu()),
        obs=mk(torch.empty(N, 934)), rew=mk(torch.empty(N)), raw=mk(torch.empty(N, 4)),
        reset=mk(torch.empty(N, dtype=torch.uint8)), term=mk(torch.empty(N, dtype=torch.uint8)),
    )  # fmt: skip
    args = _cabi.PhcHostStepArgs(*[h[k].data_ptr() for k in ("state", "prog", "start", "off", "goff", "ids", "obs",
                                                             "rew", "raw", "reset", "term")])  # fmt: skip
    # auto / tuned: every candidate schedule is timed over the first calls, then one is kept
    calls = capi.phc_host_step_tuning_calls(ctx) + 2 if mode in ("pinned_auto", "pinned_tuned") else 1
    assert calls == {"pinned_auto": 2 + 2 * 5 + 2, "pinned_tuned": 2 + 6 * 5 + 2}.get(mode, 1)
    for i in range(calls):
        h["prog"].copy_(clock.progress_buf.cpu())
        h["obs"].fill_(float("nan"))
        _cabi.check(capi.phc_host_step(ctx, C.byref(args), N), "host step")
        assert torch.equal(h["obs"], env.obs_buf.cpu()), f"call {i}"
    assert capi.phc_host_step_h2d_bytes(ctx, N) == N * (1248 + 2 + 4 + 4 + 12 + 8)
    path = capi.phc_host_step_path(ctx)
    assert path == {"pinned_direct": 1, "pinned_staged": 2, "pageable_staged": 2}.get(mode, path) and path in (1, 2)
    assert capi.phc_host_step_chunks(ctx) in ((2, 3, 4, 6) if mode == "pinned_tuned" else (3,))
    capi.phc_host_step_destroy(ctx)
    assert torch.equal(h["obs"], env.obs_buf.cpu()) and torch.equal(h["rew"], env.rew_buf.cpu())
    assert torch.equal(h["raw"], env.reward_raw[:, :4].cpu())
    assert torch.equal(h["reset"].bool(), env.reset_buf.cpu()) and torch.equal(h["term"].bool(), env._terminate_buf.cpu())
    assert torch.equal(h["prog"], env.progress_buf.cpu())
    if pinned:  # device pointers are refused, not dereferenced by the CPU
        bad = _cabi.PhcHostStepArgs(*[h[k].data_ptr() for k in ("state", "prog", "start", "off", "goff", "ids", "obs",
                                                                "rew", "raw", "reset", "term")])  # fmt: skip
        bad.state = env._rigid_body_state_reshaped.data_ptr()
        ctx2 = C.c_void_p()
        _cabi.check(capi.phc_host_step_create(lib.handle, N, 1, 3, term, 0xFFFFFF, 0, 1, synth.SIM_DT, C.byref(spec),
                                              C.byref(ctx2)), "create")  # fmt: skip
        assert capi.phc_host_step(ctx2, C.byref(bad), N) == -4  # PHC_ERR_UNSUPPORTED
        capi.phc_host_step_destroy(ctx2)


@pytest.mark.parametrize("regime", ["aligned30", "mixed_fps_unaligned"])
def test_speculation_miss_paths_do_not_change_results(regime):
    """The fast kernel speculates on the clock before its stream dependency resolves.  Perturbing
    what it speculated on (progress behind by 1 or 2, start time, motion id) must leave every
    output bit-identical: the validation selects other candidates or redoes blends and copies."""
    from humanoid_b200 import HumanoidPHC, _cabi

    kw = dict(max_frames=90, max_progress=40)
    if regime != "aligned30":
        kw.update(ids="random", aligned=False, fps_choices=(30, 60, 120), min_frames=60, max_frames=400)
    N = 3001
    lib_data, clock, state = _gpu_case(N, 257, 209, **kw)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    outs = []
    try:
        for fault in (0, 1, 2, 4, 8, 3):
            assert capi.phc_set_option(_cabi.OPT_TEST_SPEC_FAULT, fault) == 0
            env = HumanoidPHC(lib, N, device=DEV)
            env.set_sim_state(state)
            env.set_clock(clock)
            for _ in range(3):
                env.step()
            torch.cuda.synchronize()
            outs.append(env)
    finally:
        capi.phc_set_option(_cabi.OPT_TEST_SPEC_FAULT, 0)
    g = outs[0]
    for f in outs[1:]:
        assert torch.equal(f.obs_buf, g.obs_buf)
        assert torch.equal(f.rew_buf, g.rew_buf) and torch.equal(f.reward_raw, g.reward_raw)
        assert torch.equal(f.reset_buf, g.reset_buf) and torch.equal(f._terminate_buf, g._terminate_buf)
        assert torch.equal(f.progress_buf, g.progress_buf)


def test_library_rewritten_in_place_is_seen_by_the_next_step():
    """The fast kernel reads clip metadata and frame rows before its dependency wait.  Rewriting the library in
    place + repack() right before a step, on the same stream and without a sync, must still be seen: the first step
    after phc_lib_pack launches without the speculation (and releases its dependents only after its own wait)."""
    N = 4096
    lib_data, clock, state = _gpu_case(N, N, seed=77, max_progress=30)
    env = env_from(lib_data, clock, state)
    for _ in range(3):  # steady state: speculating launches
        env.progress_buf.copy_(clock.progress_buf)
        env.step()
    lib = env._motion_lib
    env.progress_buf.copy_(clock.progress_buf)
    # a long kernel in front, so that the rewrite is still in flight when the step's blocks become resident
    big = torch.empty(1 << 26, dtype=torch.float32, device=DEV)
    big.fill_(1.0)
    lib.gts.add_(0.125)
    lib.gvs.mul_(2.0)
    lib.repack()
    env.step()
    env.step()  # the second step after the rewrite speculates again
    got = [t.clone() for t in (env.obs_buf, env.rew_buf, env.reset_buf, env._terminate_buf)]
    torch.cuda.synchronize()
    # the same two steps on a fresh env over a fresh library built from the rewritten tensors
    d2 = lib_data.as_dict()
    d2["gts"], d2["gvs"] = lib.gts.clone(), lib.gvs.clone()
    env2 = env_from(synth.MotionData(**d2), clock, state)
    torch.cuda.synchronize()
    env2.step()
    env2.step()
    for a, b, nm in zip(got, (env2.obs_buf, env2.rew_buf, env2.reset_buf, env2._terminate_buf), ("obs", "rew", "reset", "term")):
        assert torch.equal(a, b), f"{nm}: the step after an in-place library rewrite used stale frames"


def test_plain_attribute_assignment_is_honoured_by_the_fused_step():
    """Reference-style scripts flip ``env.flag_im_eval`` / edit ``env.rwd_specs`` / ``enable_early_termination`` by
    plain assignment; the cached step arguments must not hide that (the unfused path reads them every call)."""
    lib_data, clock, state = _gpu_case(1024, 64, seed=91, max_progress=30)
    env, ref = env_from(lib_data, clock, state), env_from(lib_data, clock, state)
    env.step()  # builds and caches the step arguments
    for e in (env, ref):
        e.progress_buf.copy_(clock.progress_buf)
        e.flag_im_eval = True
        e.rwd_specs["k_pos"] = 50.0
        e.rwd_specs["w_rot"] = 0.2
        e.set_termination_distances(0.2)
    env.step()
    ref.post_physics_step_unfused()
    assert torch.equal(env._terminate_buf, ref._terminate_buf) and torch.equal(env.reset_buf, ref.reset_buf)
    assert torch.equal(env.rew_buf, ref.rew_buf) and torch.equal(env.obs_buf, ref.obs_buf)
    assert env.extras["mpjpe"] is not None
    for e in (env, ref):
        e.progress_buf.copy_(clock.progress_buf)
        e.flag_im_eval = False
        e.enable_early_termination = False
    env.step()
    ref.post_physics_step_unfused()
    assert torch.equal(env._terminate_buf, ref._terminate_buf) and not bool(env._terminate_buf.any())
    assert torch.equal(env.reset_buf, ref.reset_buf)


def test_env_reset_between_steps_is_seen_by_the_next_step():
    """Clock arrays rewritten on the device between two steps (what a reset does) with no host sync."""
    lib_data, clock, state = make_case_cpu(num_envs=640, num_motions=40, seed=111, max_progress=30)
    env = env_from(lib_data, clock, state)
    ol = O.OracleMotionLib(lib_data)
    prog = clock.progress_buf.clone()
    term = torch.full((24,), 0.25)
    args = lambda c, p: (ol, state, p, c.motion_start_times, c.motion_start_times_offset, c.global_offset,  # noqa: E731
                         c.sampled_motion_ids, term, synth.SIM_DT)
    O.step(*args(clock, prog))
    env.step()
    # "reset" every third env: new clip, new start time, progress 0 — device-side writes, same stream
    sel = torch.arange(0, 640, 3)
    c2 = clock.clone()
    c2.sampled_motion_ids[sel] = (clock.sampled_motion_ids[sel] + 7) % 40
    c2.motion_start_times[sel] = (torch.arange(sel.numel()) % 5).float() * (1 / 30)
    prog[sel] = 0
    sel_d = sel.to(DEV)
    env._sampled_motion_ids[sel_d] = c2.sampled_motion_ids[sel].to(DEV)
    env._motion_start_times[sel_d] = c2.motion_start_times[sel].to(DEV)
    env.progress_buf[sel_d] = 0
    want = O.step(*args(c2, prog))
    env.step()
    assert_equal_exact(env.progress_buf, prog, "progress")
    assert_equal_exact(env.reset_buf, want[3], "reset")
    assert_equal_exact(env._terminate_buf, want[4], "terminated")
    assert_close(env.obs_buf, want[0], what="obs after reset", **OBS_TOL)
    assert_close(env.rew_buf, want[1], what="reward after reset", **OBS_TOL)


def test_amass_scale_library_random_queries():
    """BASELINE config 4 at scale: ~0.6 M frames, mixed fps, long clips, random ids and unaligned
    times.  Size-independent checks: fused == per-function kernels bitwise, frame indices in range,
    and a 512-env sample against the oracle (which only gathers the frames it needs)."""
    from humanoid_b200 import HumanoidPHC

    M, N = 1500, 32768
    gen = torch.Generator(device=DEV).manual_seed(5)
    nf = torch.randint(60, 760, (M,), generator=gen, device=DEV)
    nf[:8] = 7000  # a few very long clips
    lib_data = synth.make_motion_lib(M, fps_choices=(30, 60, 120), seed=301, device=DEV, frames_per_motion=nf)
    assert lib_data.total_frames > 600_000
    lib = MotionLib(lib_data, device=DEV)
    clock = synth.make_clock(lib_data, N, seed=302, ids="random", aligned=False, max_progress=60)
    t = synth.reward_time(clock, extra_steps=1)
    ref = lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset, with_frame_info=True)
    assert int(ref["frame_idx0"].min()) >= 0
    assert bool((ref["frame_idx1"] < lib_data.motion_num_frames[clock.sampled_motion_ids]).all())
    assert bool((ref["frame_idx1"] - ref["frame_idx0"]).clamp(0, 1).eq(ref["frame_idx1"] - ref["frame_idx0"]).all())
    state = synth.make_sim_state(ref, seed=303)
    a, b = HumanoidPHC(lib, N, device=DEV), HumanoidPHC(lib, N, device=DEV)
    for env in (a, b):
        env.set_sim_state(state)
        env.set_clock(clock)
    a.step()
    b.post_physics_step_unfused()
    assert torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.rew_buf, b.rew_buf)
    assert torch.equal(a.reset_buf, b.reset_buf) and torch.equal(a._terminate_buf, b._terminate_buf)
    # oracle on a sample of envs: CPU copies of only the four frame tensors the step reads
    sel = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:512]
    cpu = {k: (v.cpu() if k in ("gts", "grs", "gvs", "gavs") or v.dim() < 2 or v.shape[0] == M else v[:1].cpu())
           for k, v in lib_data.as_dict().items()}  # fmt: skip
    ol = O.OracleMotionLib(cpu)
    ol.lrs, ol.dvs, ol._motion_aa = cpu["grs"], cpu["gvs"][:, :23], torch.zeros(lib_data.total_frames, 1)
    c = synth.Clock(**{k: v.cpu()[sel] for k, v in clock.__dict__.items()})
    prog = c.progress_buf.clone()
    want = O.step(ol, state.cpu()[sel], prog, c.motion_start_times, c.motion_start_times_offset, c.global_offset,
                  c.sampled_motion_ids, torch.full((24,), 0.25), synth.SIM_DT)  # fmt: skip
    sel_d = sel.to(DEV)
    assert_equal_exact(a.reset_buf[sel_d], want[3], "reset")
    assert_equal_exact(a._terminate_buf[sel_d], want[4], "terminated")
    assert_close(a.obs_buf[sel_d], want[0], what="obs (AMASS-scale sample)", **OBS_TOL)
    assert_close(a.rew_buf[sel_d], want[1], what="reward (AMASS-scale sample)", **OBS_TOL)


def test_config5_T10_fused_equals_per_function_at_16384():
    from humanoid_b200 import HumanoidPHC

    N = 16384
    lib_data, clock, state = _gpu_case(N, 1024, 304, max_frames=120, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    a, b = HumanoidPHC(lib, N, device=DEV, time_steps=10), HumanoidPHC(lib, N, device=DEV, time_steps=10)
    for env in (a, b):
        env.set_sim_state(state)
        env.set_clock(clock)
    a.step()
    b.post_physics_step_unfused()
    assert a.obs_buf.shape == (N, 358 + 5760)
    assert torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.rew_buf, b.rew_buf)
    assert torch.equal(a.reset_buf, b.reset_buf) and torch.equal(a._terminate_buf, b._terminate_buf)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "per_function"])
def test_power_reward_vs_oracle(fused):
    """use_power_reward (humanoid_phc.py:1297-1305): -coef * sum |dof_force * dof_vel|, zero while
    progress <= 3, added to rew_buf and stored in reward_raw[:, -1]; dof_vel is the stride-2 view."""
    lib_data, clock, state = make_case_cpu(num_envs=1500, num_motions=48, seed=112, max_progress=12)
    gen = torch.Generator().manual_seed(4)
    force = torch.randn(1500, 69, generator=gen) * 40
    dvel = torch.randn(1500, 69, generator=gen) * 3
    prog = clock.progress_buf.clone()
    want = O.step(O.OracleMotionLib(lib_data), state, prog, clock.motion_start_times, clock.motion_start_times_offset,
                  clock.global_offset, clock.sampled_motion_ids, torch.full((24,), 0.25), synth.SIM_DT,
                  dof_force=force, dof_vel=dvel, rew_power_coef=0.0005)  # fmt: skip
    env = env_from(lib_data, clock, state, use_power_reward=True)
    env.dof_force_tensor.copy_(force)
    env._dof_vel.copy_(dvel)
    assert env._dof_vel.stride(1) == 2
    if fused:
        env.step()
    else:
        env.post_physics_step_unfused()
    assert (prog <= 3).any() and (prog > 3).any()
    assert_close(env.rew_buf, want[1], what="reward incl. power", **OBS_TOL)
    assert_close(env.reward_raw, want[2], what="reward_raw incl. power column", **OBS_TOL)
    assert float(env.reward_raw[:, 4].cpu()[prog <= 3].abs().max()) == 0.0


def _oracle_env_state(N, clock, state):
    return dict(
        state=state.clone(), root=torch.zeros(N, 13), dof_pos=torch.zeros(N, 69), dof_vel=torch.zeros(N, 69),
        prog=clock.progress_buf.clone(), reset=torch.ones(N, dtype=torch.bool), term=torch.ones(N, dtype=torch.bool),
        start=clock.motion_start_times.clone(), soff=clock.motion_start_times_offset.clone(),
        goff=clock.global_offset.clone(), ids=clock.sampled_motion_ids.clone(), obs=torch.zeros(N, 934),
    )  # fmt: skip


def _oracle_reset(ol, s, env_ids, phase, **kw):
    return O.reset_envs(ol, env_ids, phase, s["state"], s["root"], s["dof_pos"], s["dof_vel"], s["prog"], s["reset"],
                        s["term"], s["start"], s["soff"], s["goff"], s["ids"], s["obs"], synth.SIM_DT, **kw)  # fmt: skip


def _oracle_obs_on(ol, state_cpu, s, env_ids):
    """Observation of env_ids computed by the oracle from a GIVEN sim state (the one the device
    wrote), so that the comparison does not go through each side's own 1e-7-different root rotation
    (the heading is ill-conditioned when the root's x axis is near vertical)."""
    J = 24
    pos, rot = state_cpu[env_ids, :J, 0:3], state_cpu[env_ids, :J, 3:7]
    vel, ang = state_cpu[env_ids, :J, 7:10], state_cpu[env_ids, :J, 10:13]
    so = O.self_obs_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    t1 = (s["prog"][env_ids] + 1) * synth.SIM_DT + s["start"][env_ids] + s["soff"][env_ids]
    r = ol.get_motion_state(s["ids"][env_ids], t1, s["goff"][env_ids])
    to = O.imitation_obs_v6(pos[:, 0], rot[:, 0], pos, rot, vel, ang, r["rg_pos"], r["rb_rot"], r["body_vel"],
                            r["body_ang_vel"], 1, True)  # fmt: skip
    return torch.cat([so, to], dim=-1)


def test_sample_time_interval_vs_reference_fixture(golden):
    """The start times a reset writes are the reference's sample_time_interval for the same phase."""
    from humanoid_b200 import HumanoidPHC

    g = golden("sample_time")
    lens, ids, phase = g.inp("motion_lengths"), g.inp("ids"), g.inp("phase")
    n, M = ids.numel(), lens.numel()
    nf = ((lens * 30).round().long() + 1).clamp_min(2)
    lib_data = synth.make_motion_lib(M, seed=5, frames_per_motion=nf)
    lib_data.motion_lengths = lens.clone()  # the lengths the fixture was made with
    env = HumanoidPHC(MotionLib(lib_data, device=DEV), n, device=DEV)
    env._sampled_motion_ids.copy_(ids)
    env.reset(torch.arange(n), phase=cuda(phase))
    assert torch.equal(env._motion_start_times.cpu(), g.out("motion_time"))
    assert int(env.progress_buf.abs().sum()) == 0 and not bool(env.reset_buf.any())


@pytest.mark.parametrize("random_init", [True, False], ids=["StateInit.Random", "StateInit.Start"])
def test_reset_subset_vs_oracle(random_init):
    lib_data, clock, state = make_case_cpu(num_envs=900, num_motions=60, seed=113, max_progress=30,
                                           fps_choices=(30, 60), min_frames=40, max_frames=200)  # fmt: skip
    env = env_from(lib_data, clock, state)
    env.state_init_random = random_init
    ol = O.OracleMotionLib(lib_data)
    s = _oracle_env_state(900, clock, state)
    gen = torch.Generator().manual_seed(9)
    env_ids = torch.randperm(900, generator=gen)[:333]
    phase = torch.rand(333, generator=gen)
    before = env._rigid_body_state_reshaped.clone()
    _oracle_reset(ol, s, env_ids, phase, random_init=random_init)
    env.reset(cuda(env_ids), phase=cuda(phase))
    keep = torch.ones(900, dtype=torch.bool)
    keep[env_ids] = False
    # untouched envs are bit-identical; reset envs match the oracle
    assert torch.equal(env._rigid_body_state_reshaped.cpu()[keep], before.cpu()[keep])
    assert torch.equal(env._motion_start_times.cpu(), s["start"])
    assert torch.equal(env._motion_start_times_offset.cpu(), s["soff"]) and torch.equal(env._global_offset.cpu(), s["goff"])
    assert_equal_exact(env.progress_buf, s["prog"], "progress")
    assert_equal_exact(env.reset_buf, s["reset"], "reset_buf")
    assert_equal_exact(env._terminate_buf, s["term"], "terminate_buf")
    assert_close(env._rigid_body_state_reshaped, s["state"], what="rigid body state", **OBS_TOL)
    assert_close(env._humanoid_root_states[cuda(env_ids)], s["root"][env_ids], what="root states", **OBS_TOL)
    assert_close(env._dof_pos[cuda(env_ids)], s["dof_pos"][env_ids], what="dof_pos", **DOF_TOL)
    assert_close(env._dof_vel[cuda(env_ids)], s["dof_vel"][env_ids], what="dof_vel", **OBS_TOL)
    want_obs = _oracle_obs_on(ol, env._rigid_body_state_reshaped.cpu(), s, env_ids)
    assert_close(env.obs_buf[cuda(env_ids)], want_obs, what="obs of reset envs", **OBS_TOL)
    assert float(env.obs_buf.cpu()[keep].abs().sum()) == 0.0  # rows of other envs untouched


def test_rollout_with_device_side_resets_vs_oracle():
    """step -> reset the envs flagged by reset_buf (no host sync) -> step ..., six times."""
    N = 512
    lib_data, clock, state = make_case_cpu(num_envs=N, num_motions=N, seed=114, min_frames=12, max_frames=40, max_progress=8)
    env = env_from(lib_data, clock, state)
    ol = O.OracleMotionLib(lib_data)
    s = _oracle_env_state(N, clock, state)
    term = torch.full((24,), 0.25)
    gen = torch.Generator().manual_seed(10)
    for it in range(6):
        want = O.step(ol, s["state"], s["prog"], s["start"], s["soff"], s["goff"], s["ids"], term, synth.SIM_DT)
        s["reset"], s["term"] = want[3].clone(), want[4].clone()
        env.step()
        assert_equal_exact(env.reset_buf, want[3], f"reset @{it}")
        assert_close(env.obs_buf, want[0], what=f"obs @{it}", **OBS_TOL)
        assert_close(env.rew_buf, want[1], what=f"reward @{it}", **OBS_TOL)
        phase = torch.rand(N, generator=gen)
        ids = torch.nonzero(s["reset"]).squeeze(-1)  # the oracle side may sync; the device side does not
        assert 0 < ids.numel() < N
        s["obs"] = want[0].clone()
        _oracle_reset(ol, s, ids, phase[ids])
        env.reset_done(cuda(phase))
        assert_equal_exact(env.progress_buf, s["prog"], f"progress after reset @{it}")
        assert torch.equal(env._motion_start_times.cpu(), s["start"])
        # the "physics" of this test: the state the reset wrote is what the next step sees
        assert_close(env._rigid_body_state_reshaped, s["state"], what=f"state after reset @{it}", **OBS_TOL)
        want_obs = want[0].clone()
        want_obs[ids] = _oracle_obs_on(ol, env._rigid_body_state_reshaped.cpu(), s, ids)
        assert_close(env.obs_buf, want_obs, what=f"obs after reset @{it}", **OBS_TOL)
        s["state"] = env._rigid_body_state_reshaped.cpu().clone()  # keep both sides on identical inputs


@pytest.mark.parametrize("N,flag,T", [(1029, "some", 1), (515, "none", 1), (64, "all", 3), (7, "some", 10), (300, "some", 2)])
def test_reset_with_reset_buf_as_its_own_mask_equals_the_copied_mask_bitwise(N, flag, T):
    """phc_reset_envs reads each mask byte once before it clears reset_buf, so the flags can select the envs
    directly (reset_done without AMP); same buffers, bit for bit, as with a copy of the flags, and with nothing
    flagged nothing is written at all."""
    lib_data, clock, state = make_case_cpu(num_envs=N, num_motions=max(4, N // 3), seed=131, max_progress=30,
                                           fps_choices=(30, 60), min_frames=40, max_frames=200)  # fmt: skip
    a, b = env_from(lib_data, clock, state, time_steps=T), env_from(lib_data, clock, state, time_steps=T)
    gen = torch.Generator().manual_seed(11)
    flags = torch.rand(N, generator=gen) < 0.3 if flag == "some" else torch.full((N,), flag == "all")
    phase = cuda(torch.rand(N, generator=gen))
    for env in (a, b):
        env.step()
        env.reset_buf.copy_(cuda(flags))
        env._terminate_buf.copy_(cuda(flags))
    before = {k: getattr(a, k).clone() for k in ("obs_buf", "_rigid_body_state_reshaped", "progress_buf")}
    a._reset_masked(a.reset_buf.clone(), phase)  # a copy of the flags selects the envs
    b.reset_done(phase)  # reset_buf itself does
    assert b.use_amp_obs is False
    for k in ("obs_buf", "_rigid_body_state_reshaped", "_humanoid_root_states", "_dof_state", "progress_buf", "reset_buf",
              "_terminate_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset"):  # fmt: skip
        if hasattr(a, k):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert not bool(b.reset_buf.any()) and not bool(b._terminate_buf[cuda(flags)].any())
    if flag == "none":
        for k, v in before.items():
            assert torch.equal(getattr(b, k), v), k
    else:
        assert not torch.equal(b.obs_buf, before["obs_buf"])
        # the rows the reset kernel wrote are, bit for bit, what the step kernel computes from the new state and clock
        rows = b.obs_buf[cuda(flags)].clone()
        b.post_physics_step(advance_progress=False)
        assert torch.equal(b.obs_buf[cuda(flags)], rows)


@pytest.mark.parametrize("N,flag", [(1029, "some"), (515, "none"), (37, "all"), (5, "some")])
def test_amp_init_with_slot0_in_the_same_launch_equals_the_two_launches_bitwise(N, flag):
    """_init_amp_obs(env_ids) (:791-799) as one launch (PhcAmpEnvArgs.init_slot0) against
    _compute_amp_observations(env_ids) + _init_amp_obs_ref(env_ids) as two; untouched envs keep their rows."""
    import ctypes as C

    from humanoid_b200 import _cabi

    lib_data, clock, state = make_case_cpu(num_envs=N, num_motions=max(4, N // 3), seed=137, max_progress=30,
                                           fps_choices=(30, 60), min_frames=40, max_frames=200)  # fmt: skip
    a, b = env_from(lib_data, clock, state, use_amp_obs=True), env_from(lib_data, clock, state, use_amp_obs=True)
    gen = torch.Generator().manual_seed(12)
    flags = cuda(torch.rand(N, generator=gen) < 0.3 if flag == "some" else torch.full((N,), flag == "all"))
    for env in (a, b):
        env.step()
        env._amp_obs_buf.copy_(cuda(torch.rand(env._amp_obs_buf.shape, generator=torch.Generator().manual_seed(13))))
        env._amp_obs_demo_buf.fill_(7.0)
    before = b._amp_obs_buf.clone()
    a._amp_step(roll=False, mask=flags)
    args, keep = a._amp_args(flags)
    _cabi.check(_cabi.load().phc_amp_init_ref(a._motion_lib.handle, C.byref(args), a._sampled_motion_ids.data_ptr(),
                                              a._motion_start_times.data_ptr(), a.dt, N, _cabi.stream_ptr(a.device)),
                "phc_amp_init_ref")  # fmt: skip
    b._init_amp_obs_masked(flags)
    assert torch.equal(a._amp_obs_buf, b._amp_obs_buf)
    assert torch.equal(a._amp_obs_demo_buf, b._amp_obs_demo_buf)
    assert torch.equal(b._amp_obs_buf[~flags], before[~flags])
    assert bool((b._amp_obs_demo_buf[~flags] == 7.0).all())
    if flag != "none":
        assert not torch.equal(b._amp_obs_buf[flags], before[flags])
        assert torch.equal(b._amp_obs_demo_buf[flags], b._amp_obs_buf[flags])


@pytest.mark.parametrize("groups", [2, 1], ids=["two_groups", "one_group"])
@pytest.mark.parametrize("N,T", [(1027, 10), (64, 16), (515, 2), (2050, 3), (9, 5), (9473, 10), (12001, 3)])
def test_multi_T_kernel_equals_generic_kernel_bitwise(N, T, groups):
    """T > 1 dispatches to the pipelined TMA kernels (two query groups per block by default, one as the cross-check);
    the generic kernel is the same math, three steps deep, single-wave and multi-wave grids.  Also a regression test:
    the clock must be advanced only after every query lane has read it."""
    from humanoid_b200 import HumanoidPHC, _cabi

    lib_data, clock, state = _gpu_case(N, 64, 210, max_frames=60, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    outs = []
    try:
        assert capi.phc_set_option(_cabi.OPT_MULTI_GROUPS, groups) == 0
        for generic in (1, 0):
            assert capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, generic) == 0
            env = HumanoidPHC(lib, N, device=DEV, time_steps=T, obs_moments=True, use_power_reward=True)
            env.set_sim_state(state)
            env.set_clock(clock)
            env.dof_force_tensor.normal_(generator=torch.Generator(device=DEV).manual_seed(1))
            env._dof_vel.copy_(torch.randn(N, 69, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV))
            snaps = []
            for _ in range(3):
                env.step()
                snaps.append([getattr(env, k).clone() for k in ("obs_buf", "rew_buf", "reward_raw", "reset_buf",
                                                                "_terminate_buf", "progress_buf")])
            torch.cuda.synchronize()
            outs.append((env, snaps))
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
        capi.phc_set_option(_cabi.OPT_MULTI_GROUPS, 0)
    (g, gs), (f, fs) = outs
    for k, (a, b) in enumerate(zip(fs, gs)):
        for x, y, nm in zip(a, b, ("obs", "rew", "reward_raw", "reset", "terminate", "progress")):
            assert torch.equal(x, y), f"step {k}: {nm}"
    torch.testing.assert_close(f.obs_moments, g.obs_moments, rtol=1e-12, atol=1e-9)


def test_amp_observations_vs_reference_fixture(golden):
    from humanoid_b200 import build_amp_observations_smpl, dof_subset_smpl

    g = golden("amp_obs")
    a = {k: cuda(g.inp(k)) for k in ("root_pos", "root_rot", "root_vel", "root_ang_vel", "dof_pos", "dof_vel",
                                      "key_body_pos", "shape", "limb", "dof_subset")}  # fmt: skip
    assert torch.equal(dof_subset_smpl(DEV), a["dof_subset"])
    # the env hands the stride-2 view of the interleaved dof state (humanoid_phc.py:535-536)
    n = a["dof_pos"].shape[0]
    dof_state = torch.zeros(n, 69, 2, device=DEV)
    dof_state[..., 0], dof_state[..., 1] = a["dof_pos"], a["dof_vel"]
    for name in ("default", "all_dofs", "global_root_no_height", "not_upright_with_params"):
        fl = [bool(x) for x in g.inp(f"flags.{name}")]
        o = build_amp_observations_smpl(a["root_pos"], a["root_rot"], a["root_vel"], a["root_ang_vel"],
                                        dof_state[..., 0], dof_state[..., 1], a["key_body_pos"], a["shape"], a["limb"],
                                        a["dof_subset"], *fl)  # fmt: skip
        assert_close(o, g.out(name), what=f"amp obs {name}", **OBS_TOL)


@pytest.mark.parametrize("res_action", [False, True])
def test_action_to_pd_targets_bit_exact(res_action):
    from humanoid_b200 import HumanoidPHC

    lib = MotionLib(synth.make_motion_lib(4, 6, 9), device=DEV)
    n = 777
    env = HumanoidPHC(lib, n, device=DEV)
    g = torch.Generator().manual_seed(6)
    act = torch.randn(n, 69, generator=g).clamp(-1, 1)
    off, sc = torch.randn(69, generator=g), torch.rand(69, generator=g) * 3
    refp, dofp = torch.randn(n, 69, generator=g), torch.randn(n, 69, generator=g)
    env._pd_action_offset.copy_(off)
    env._pd_action_scale.copy_(sc)
    env._dof_pos.copy_(dofp)
    got = env._action_to_pd_targets(cuda(act), res_action=res_action, ref_dof_pos=cuda(refp), freeze_hand=True, freeze_toe=True)
    want = O.action_to_pd_targets(act, off, sc, res_action, refp, dofp, zero_joints=(3, 7, 17, 22))
    assert torch.equal(got.cpu(), want)  # one multiply, one add, min/max: no room for rounding differences
    assert float(got[:, 9:12].abs().sum()) == 0.0 and float(got[:, 66:69].abs().sum()) == 0.0


# ---------------------------------------------------------------------------------------
# episode bookkeeping of the pufferlib wrapper (clean_pufferl/env.py:121-159)
# ---------------------------------------------------------------------------------------
def _episode_state(n, cols, device):
    return dict(
        terminals=torch.zeros(n, dtype=torch.bool, device=device), truncations=torch.zeros(n, dtype=torch.bool, device=device),
        masks=torch.ones(n, dtype=torch.bool, device=device), episode_returns=torch.zeros(n, device=device),
        episode_lengths=torch.zeros(n, dtype=torch.int32, device=device), raw_rewards=torch.zeros(cols, device=device),
        stats=torch.zeros(4, dtype=torch.float64, device=device), workspace=torch.zeros(16, dtype=torch.float64, device=device),
    )  # fmt: skip


def _episode_update_cabi(st, reset, terminate, rewards, raw):
    from humanoid_b200 import _cabi

    reset, terminate, rewards, raw = cuda(reset), cuda(terminate), cuda(rewards), cuda(raw).contiguous()
    a = _cabi.PhcEpisodeArgs()
    a.reset, a.terminate, a.rewards = reset.data_ptr(), terminate.data_ptr(), rewards.data_ptr()
    a.reward_raw, a.reward_raw_stride, a.reward_raw_cols = raw.data_ptr(), raw.stride(0), raw.shape[1]
    a.terminals, a.truncations, a.masks = st["terminals"].data_ptr(), st["truncations"].data_ptr(), st["masks"].data_ptr()
    a.episode_returns, a.episode_lengths = st["episode_returns"].data_ptr(), st["episode_lengths"].data_ptr()
    a.stats, a.raw_rewards, a.workspace = st["stats"].data_ptr(), st["raw_rewards"].data_ptr(), st["workspace"].data_ptr()
    _cabi.check(_cabi.load().phc_episode_update(a, reset.shape[0], _cabi.stream_ptr(DEV)), "phc_episode_update")
    torch.cuda.synchronize()
    assert not st["workspace"].any(), "workspace must be left zero"


def test_episode_bookkeeping_vs_reference_recording(golden):
    from conftest import replay_episode_golden

    replay_episode_golden(golden("episode"), lambda n, c: _episode_state(n, c, DEV), _episode_update_cabi, read=lambda t: t.cpu())


@pytest.mark.parametrize("N", [1, 257, 4096, 8192, 8193, 70001, 300000])  # one-block path up to 8192 envs, tickets beyond
def test_episode_bookkeeping_vs_oracle_many_blocks(N):
    g = torch.Generator().manual_seed(N)
    want = _episode_state(N, 5, "cpu")
    want.pop("workspace")
    got = _episode_state(N, 5, DEV)
    for k in range(4):
        rewards = torch.rand(N, generator=g)
        raw = torch.rand(N, 5, generator=g)
        reset = torch.rand(N, generator=g) < (0.3 if k else 0.0)
        terminate = reset & (torch.rand(N, generator=g) < 0.5)
        O.episode_update(want, reset, terminate, rewards, raw)
        _episode_update_cabi(got, reset, terminate, rewards, raw)
    for key in ("terminals", "truncations", "masks", "episode_lengths", "episode_returns"):
        assert_equal_exact(got[key], want[key], key)
    assert_close(got["raw_rewards"], want["raw_rewards"], rtol=1e-6, atol=1e-7, what="raw_rewards")
    assert_equal_exact(got["stats"][[0, 2, 3]], want["stats"][[0, 2, 3]], "counts")
    assert_close(got["stats"][1], want["stats"][1], rtol=1e-12, atol=0, what="sum of returns")


def test_puffer_env_step_matches_oracle():
    """PHCPufferEnv.step = fused step + bookkeeping kernel + device-side reset of the flagged envs."""
    from humanoid_b200 import PHCPufferEnv

    N = 1500
    lib_data, clock, state = make_case_cpu(num_envs=N, num_motions=40, seed=131, max_progress=30)
    env = env_from(lib_data, clock, state)
    pe = PHCPufferEnv(env, num_actions=69, log_interval=3)
    want = _episode_state(N, 5, "cpu")
    want.pop("workspace")
    g = torch.Generator().manual_seed(5)
    for k in range(3):
        actions = torch.rand(N, 69, generator=g) * 4 - 2
        phase = torch.rand(N, generator=g)
        obs, rew, term, trunc, info = pe.step(actions.numpy(), phase_by_env=cuda(phase))
        # the oracle bookkeeping on the flags / rewards the step produced (their parity is tested above)
        assert torch.equal(pe.actions.cpu(), actions.clamp(-1, 1))
        reset_k, rew_k = (term | trunc).cpu(), rew.cpu()
        O.episode_update(want, reset_k, term.cpu(), rew_k, env.reward_raw.cpu())
        assert_equal_exact(pe.episode_lengths, want["episode_lengths"], f"lengths[{k}]")
        assert_equal_exact(pe.episode_returns, want["episode_returns"], f"returns[{k}]")
        assert_equal_exact(pe.masks, want["masks"], f"masks[{k}]")
        assert not env.reset_buf.any() and (env.progress_buf[reset_k.to(DEV)] == 0).all()
    assert len(info) == 1 and set(info[0]) >= {"episode_return", "episode_length", "truncated_rate", "rew_body_pos"}
    s = want["stats"].tolist()
    assert s[0] > 0 and pe.episode_count == int(s[0])
    assert abs(info[0]["episode_return"] - s[1] / s[0]) <= 1e-9 * abs(s[1] / s[0])
    assert abs(info[0]["rew_body_pos"] - float(want["raw_rewards"][0]) / 3) < 1e-6


@pytest.mark.parametrize("N,T,mode", [(1503, 1, "fast"), (4096, 1, "fast_power_norm"), (700, 1, "generic"), (515, 3, "multi")])
def test_episode_bookkeeping_fused_into_the_step_equals_the_standalone_kernel(N, T, mode):
    """PHCPufferEnv(fused=True): the step kernel (T = 1) or one small launch after it (other kernels) does the wrapper's
    bookkeeping; buffers must equal the standalone phc_episode_update path bit for bit, the logged means to 1e-12."""
    from humanoid_b200 import HumanoidPHC, PHCPufferEnv, RunningNorm, _cabi

    lib_data, clock, state = _gpu_case(N, 40, 555, max_progress=30)
    lib = MotionLib(lib_data, device=DEV)
    capi = _cabi.load()
    gen = torch.Generator().manual_seed(8)
    steps = [(torch.rand(N, 69, generator=gen) * 4 - 2, torch.rand(N, generator=gen)) for _ in range(4)]
    outs = []
    try:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if mode == "generic" else 0)
        for fused in (False, True):
            env = HumanoidPHC(lib, N, device=DEV, time_steps=T, use_power_reward="power" in mode)
            env.set_sim_state(state)
            env.set_clock(clock)
            if "power" in mode:
                env.dof_force_tensor.normal_(generator=torch.Generator(device=DEV).manual_seed(1))
                env._dof_vel.copy_(torch.randn(N, 69, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV))
            if "norm" in mode:
                env.set_obs_normalizer(RunningNorm(env.num_obs, device=DEV))
            pe = PHCPufferEnv(env, log_interval=4, fused=fused)
            pe.episode_returns.fill_(0.25)  # as if episodes were under way
            pe.episode_lengths.fill_(7)
            infos = []
            for actions, phase in steps:
                obs, rew, term, trunc, info = pe.step(cuda(actions), phase_by_env=cuda(phase))
                infos += info
            torch.cuda.synchronize()
            outs.append((pe, env, infos))
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    (a, ea, ia), (b, eb, ib) = outs
    assert torch.equal(ea.obs_buf, eb.obs_buf) and torch.equal(ea.rew_buf, eb.rew_buf)
    for k in ("terminals", "truncations", "masks", "episode_returns", "episode_lengths", "actions"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    # fused=True also resets the flagged envs inside the step's launch: every buffer the two-launch path (step, then
    # reset_done) leaves must be there bit for bit, and the normalised rows must be those of the FINAL obs rows
    for k in ("_rigid_body_state_reshaped", "_humanoid_root_states", "_dof_state", "progress_buf", "reset_buf",
              "_terminate_buf", "_motion_start_times", "_motion_start_times_offset", "_global_offset", "reward_raw"):
        assert torch.equal(getattr(ea, k), getattr(eb, k)), k
    assert not bool(eb.reset_buf.any()) and not bool(eb._terminate_buf.any())
    assert torch.equal(eb.extras["terminate"], b.terminals) and torch.equal(eb.extras["reset"], b.terminals | b.truncations)
    if "norm" in mode:
        assert torch.equal(ea.obs_norm_buf, eb.obs_norm_buf)
        assert_close(eb.obs_norm_buf, eb.obs_normalizer(eb.obs_buf), what="obs_norm_buf after resets", rtol=1e-6, atol=1e-6)
    assert len(ia) == len(ib) == 1 and a.episode_count == b.episode_count > 0
    for k, v in ia[0].items():
        assert abs(v - ib[0][k]) <= 1e-12 + 1e-6 * abs(v), (k, v, ib[0][k])  # raw_rewards accumulate in fp32 per step vs per log
    assert float(b._ep_sums.abs().max()) == 0.0  # folded and cleared


def _reset_case(N, T=1, seed=777, **env_kw):
    lib_data, clock, state = _gpu_case(N, 48, seed, max_progress=30, max_frames=90)
    lib = MotionLib(lib_data, device=DEV)
    from humanoid_b200 import HumanoidPHC

    def make():
        env = HumanoidPHC(lib, N, device=DEV, time_steps=T, **env_kw)
        env.set_sim_state(state)
        env.set_clock(clock)
        # a dof state / root state that the reset must overwrite for the flagged envs only
        env._dof_state.copy_(torch.randn(env._dof_state.shape, generator=torch.Generator().manual_seed(3)).to(DEV))
        env._humanoid_root_states.fill_(0.5)
        return env

    return make


ENV_BUFFERS = ("obs_buf", "rew_buf", "reward_raw", "_rigid_body_state_reshaped", "_humanoid_root_states", "_dof_state",
               "progress_buf", "reset_buf", "_terminate_buf", "_motion_start_times", "_motion_start_times_offset",
               "_global_offset")  # fmt: skip


@pytest.mark.parametrize("N,T,mode", [(4096, 1, "fast"), (1029, 1, "fast"), (3, 1, "fast"), (1029, 1, "fast_eval"),
                                      (1000, 1, "generic"), (515, 3, "multi"), (700, 1, "fast_start"),
                                      (1029, 1, "fast_res_action"), (515, 2, "multi_res_action")])  # fmt: skip
def test_reset_inside_the_step_equals_step_then_reset_bitwise(N, T, mode):
    """PhcStepArgs.auto_reset: the envs a step flags are re-posed, their clocks restarted and their observation rows
    recomputed inside the step's own launch.  Every buffer equals step() followed by reset_done(phase) — three steps
    deep, so that the steps after a reset (progress 0, new start time, zeroed offsets) are covered as well."""
    from humanoid_b200 import _cabi

    kw = dict(res_action=True) if "res_action" in mode else {}
    make = _reset_case(N, T, **kw)
    a, b = make(), make()
    b.enable_auto_reset(True)
    for e in (a, b):
        if "eval" in mode:
            e.flag_im_eval = True
            e.set_termination_distances(0.1)
        if "start" in mode:
            e.state_init_random = False
    gen = torch.Generator().manual_seed(4)
    capi = _cabi.load()
    capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if mode == "generic" else 0)
    try:
        for k in range(3):
            phase = cuda(torch.rand(N, generator=gen))
            a.step()
            flags = (a.reset_buf.clone(), a._terminate_buf.clone())
            a.reset_done(phase)
            b.step(phase_by_env=phase)
            torch.cuda.synchronize()
            assert 0 < int(flags[0].sum()) < N or N < 8, "the case should flag some envs and spare others"
            for name in ENV_BUFFERS:
                assert torch.equal(getattr(a, name), getattr(b, name)), f"step {k}: {name}"
            assert torch.equal(b.extras["reset"], flags[0]) and torch.equal(b.extras["terminate"], flags[1])
            assert not bool(b.reset_buf.any()) and not bool(b._terminate_buf.any())
            if "res_action" in mode:
                assert torch.equal(a.ref_dof_pos, b.ref_dof_pos), f"step {k}: ref_dof_pos"
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("auto", [True, False], ids=["reset_in_step", "reset_done"])
def test_normalised_rows_and_moments_follow_the_rows_a_reset_rewrites(auto, dtype):
    """The policy reads obs_norm_buf and RunningNorm.update must see what the policy saw: after a step whose flagged
    envs were reset (inside the step, or by reset_done()), obs_norm_buf == rn(obs_buf) for the FINAL rows and the
    fp64 partials are the column moments of the final rows — the initial reset()'s rows included."""
    from humanoid_b200 import HumanoidPHC, RunningNorm

    N = 2052
    lib_data, clock, state = _gpu_case(N, 48, 991, max_progress=30, max_frames=90)
    lib = MotionLib(lib_data, device=DEV)
    env = HumanoidPHC(lib, N, device=DEV, obs_moments=True)
    rn = RunningNorm(env.num_obs, device=DEV)
    rn.running_mean.normal_(generator=torch.Generator(device=DEV).manual_seed(1))
    rn.running_var.uniform_(0.5, 2.0, generator=torch.Generator(device=DEV).manual_seed(2))
    env.set_obs_normalizer(rn, dtype=dtype)
    env.enable_auto_reset(auto)
    gen = torch.Generator().manual_seed(6)
    env.reset(phase=cuda(torch.rand(N, generator=gen)))  # the rows the policy sees first
    want = torch.zeros(2 * env.num_obs, dtype=torch.float64, device=DEV)
    rows = 0

    def check(tag):
        nonlocal rows
        x = env.obs_buf.double()
        want[: env.num_obs] += x.sum(0)
        want[env.num_obs :] += (x * x).sum(0)
        rows += N
        ref = rn(env.obs_buf)
        if dtype == torch.bfloat16:
            assert torch.equal(env.obs_norm_buf, ref.to(torch.bfloat16)) or (
                (env.obs_norm_buf.float() - ref).abs().max() <= 0.0626), tag  # 1 bf16 ulp at |x| <= 10 (the quotient is within 1 fp32 ulp of IEEE division)
        else:
            assert_close(env.obs_norm_buf, ref, what=f"obs_norm_buf {tag}", rtol=1e-6, atol=1e-6)

    check("after reset()")
    for k in range(3):
        env.set_sim_state(state)  # the "physics": every step sees the noisy state again, so envs keep getting flagged
        phase = cuda(torch.rand(N, generator=gen))
        if auto:
            env.step(phase_by_env=phase)
        else:
            env.step()
            assert 0 < int(env.reset_buf.sum()) < N
            env.reset_done(phase)
        check(f"after step {k}")
    sums, got_rows = env.take_obs_moments()
    assert got_rows == rows
    scale = want.abs().clamp_min(1.0)
    assert float(((sums - want).abs() / scale).max()) < 1e-11, "moments must be those of the rows the policy saw"


FLAG_SETS = [(True, True, True), (False, True, True), (True, False, True), (True, True, False), (False, False, False)]


@pytest.mark.parametrize("flags", FLAG_SETS, ids=lambda f: "local%d_height%d_upright%d" % tuple(map(int, f)))
@pytest.mark.parametrize("kernel", ["fast", "generic", "reset_in_step"])
def test_self_obs_flag_variants_on_the_fused_path(flags, kernel):
    """local_root_obs / root_height_obs / has_upright_start (humanoid_phc.py:963-998, config.py:61-69) on the fused
    kernels: rows equal the per-function kernels bit for bit (those are pinned by the reference's `flags` fixture)
    and the oracle within the float tolerance; without the height column a row has 933 floats."""
    from humanoid_b200 import _cabi

    local, height, upright = flags
    N = 1030
    make = _reset_case(N, 1, seed=313, local_root_obs=local, root_height_obs=height, has_upright_start=upright)
    env, ref = make(), make()
    assert env.num_obs == 934 - (0 if height else 1)
    capi = _cabi.load()
    capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if kernel == "generic" else 0)
    try:
        if kernel == "reset_in_step":
            env.enable_auto_reset(True)
            phase = cuda(torch.rand(N, generator=torch.Generator().manual_seed(1)))
            env.step(phase_by_env=phase)
            ref.step()
            ref.reset_done(phase)
        else:
            env.step()
            ref.post_physics_step_unfused()
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    assert torch.equal(env.obs_buf, ref.obs_buf), "fused rows differ from the per-function / two-launch rows"
    assert torch.equal(env.rew_buf, ref.rew_buf) and torch.equal(env.progress_buf, ref.progress_buf)
    if kernel != "reset_in_step":
        st = env._rigid_body_state_reshaped.cpu()
        pos, rot, vel, ang = synth.body_views(st)
        want = O.self_obs_smpl_max(pos, rot, vel, ang, None, None, local, height, upright, False, False)
        # the heading is atan2(ry, rx) of the rotated x axis: where that axis is nearly vertical (rx^2 + ry^2 small, which
        # remove_base_rot of a random root rotation produces) the reference's own heading is ill-conditioned and a 1-ulp
        # difference in atan2f shows at 1e-4; those envs are compared against the per-function kernels only (above)
        rr = rot[:, 0] if upright else O.remove_base_rot(rot[:, 0])
        x_axis = O.rotate(rr, torch.tensor([1.0, 0.0, 0.0]).expand(N, 3))
        ok = (x_axis[:, 0] ** 2 + x_axis[:, 1] ** 2) > 0.05
        assert ok.float().mean() > 0.9
        assert_close(env.obs_buf[:, : want.shape[1]].cpu()[ok], want[ok], what="self obs vs oracle", **OBS_TOL)


def test_ref_dof_pos_of_the_fused_step_equals_the_motion_query():
    """res_action (humanoid_phc.py:1115-1120): the step keeps dof_pos of the query at t + dt; equal to K1's dof_pos."""
    N = 2050
    for T, generic in ((1, False), (1, True), (3, False)):
        from humanoid_b200 import _cabi

        make = _reset_case(N, T, seed=99, res_action=True)
        env = make()
        _cabi.load().phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if generic else 0)
        try:
            env.step()
        finally:
            _cabi.load().phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
        t1 = env._motion_times(plus=1)
        want = env._motion_lib.get_motion_state(env._sampled_motion_ids, t1, env._global_offset)["dof_pos"]
        assert torch.equal(env.ref_dof_pos, want), f"T={T} generic={generic}"
        # and the PD targets of the next pre-physics step use it (clamped residual form)
        act = cuda(torch.rand(N, 69, generator=torch.Generator().manual_seed(2)) * 4 - 2)
        pd = env.pre_physics_step(act, clip=1.0, actions_out=torch.empty_like(act))
        want_pd = O.action_to_pd_targets(act.clamp(-1, 1).cpu(), env._pd_action_offset.cpu(), env._pd_action_scale.cpu(),
                                         res_action=True, ref_dof_pos=want.cpu(), dof_pos=env._dof_pos.cpu())
        assert_equal_exact(pd, want_pd, "pd targets (res_action, clipped actions)")


# ---------------------------------------------------------------------------------------
# RunningNorm.forward fused into the step's obs epilogue (policies/running_norm.py:15-20)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,T,generic", [(4099, 1, False), (1000, 1, True), (600, 3, False)],
                         ids=["fast_T1", "generic_T1", "generic_T3"])  # fmt: skip
def test_fused_obs_normalisation_vs_oracle(N, T, generic):
    from humanoid_b200 import HumanoidPHC, RunningNorm, _cabi

    lib_data, clock, state = _gpu_case(N, 64, 209, max_progress=40)
    env = HumanoidPHC(MotionLib(lib_data, device=DEV), N, device=DEV, time_steps=T)
    env.set_sim_state(state)
    env.set_clock(clock)
    rn = RunningNorm(env.num_obs, device=DEV)
    g = torch.Generator().manual_seed(3)
    rn.running_mean.copy_(torch.randn(1, env.num_obs, generator=g) * 0.3)
    rn.running_var.copy_(torch.rand(1, env.num_obs, generator=g) * 2 + 1e-3)
    rn.running_var[0, :5] = 0.0  # a column that never varied: only epsilon under the root
    env.set_obs_normalizer(rn)
    capi = _cabi.load()
    try:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if generic else 0)
        env.step()
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    plain = HumanoidPHC(env._motion_lib, N, device=DEV, time_steps=T)
    plain.set_sim_state(state)
    plain.set_clock(clock)
    plain.step()
    assert torch.equal(env.obs_buf, plain.obs_buf), "the raw rows must not change"
    assert torch.equal(env.rew_buf, plain.rew_buf) and torch.equal(env.reset_buf, plain.reset_buf)
    want = O.running_norm_forward(rn.running_mean.cpu(), rn.running_var.cpu(), env.obs_buf.cpu(), rn.epsilon, rn.clip)
    assert_close(env.obs_norm_buf, want, rtol=1e-5, atol=1e-6, what="fused normalised obs")
    assert_close(env.obs_norm_buf, rn(env.obs_buf), rtol=2e-7, atol=1e-7, what="fused vs standalone forward kernel")
    assert float(env.obs_norm_buf.abs().max()) == rn.clip  # the clamp is exercised
    # an update between steps is picked up by the next launch (buffers are read at launch time)
    rn.running_mean.add_(0.5)
    env.step()
    want = O.running_norm_forward(rn.running_mean.cpu(), rn.running_var.cpu(), env.obs_buf.cpu(), rn.epsilon, rn.clip)
    assert_close(env.obs_norm_buf, want, rtol=1e-5, atol=1e-6, what="fused normalised obs, second step")


@pytest.mark.parametrize("N,T,generic", [(4099, 1, False), (777, 1, True), (300, 3, False)],
                         ids=["fast_T1", "generic_T1", "generic_T3"])  # fmt: skip
def test_fused_obs_normalisation_bf16_rows(N, T, generic):
    """PHC_STEP_OBS_NORM_BF16: the bf16 rows are the fp32 rows of the same kernel rounded to nearest-even."""
    from humanoid_b200 import HumanoidPHC, RunningNorm, _cabi

    lib_data, clock, state = _gpu_case(N, 64, 209, max_progress=40)
    lib = MotionLib(lib_data, device=DEV)
    rn = RunningNorm(358 + 576 * T, device=DEV)
    g = torch.Generator().manual_seed(5)
    rn.running_mean.copy_(torch.randn(1, rn.shape, generator=g) * 0.3)
    rn.running_var.copy_(torch.rand(1, rn.shape, generator=g) * 2 + 1e-3)
    envs = {}
    capi = _cabi.load()
    for dtype in (torch.float32, torch.bfloat16):
        env = HumanoidPHC(lib, N, device=DEV, time_steps=T)
        env.set_sim_state(state)
        env.set_clock(clock)
        env.set_obs_normalizer(rn, dtype=dtype)
        try:
            capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if generic else 0)
            env.step()
        finally:
            capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
        envs[dtype] = env
    f32, b16 = envs[torch.float32], envs[torch.bfloat16]
    assert b16.obs_norm_buf.dtype == torch.bfloat16 and b16.obs_norm_buf.shape == f32.obs_norm_buf.shape
    assert torch.equal(b16.obs_buf, f32.obs_buf) and torch.equal(b16.rew_buf, f32.rew_buf)
    assert torch.equal(b16.obs_norm_buf, f32.obs_norm_buf.to(torch.bfloat16)), "bf16 rows = RN-even rounding of the fp32 rows"
    want = O.running_norm_forward(rn.running_mean.cpu(), rn.running_var.cpu(), f32.obs_buf.cpu(), rn.epsilon, rn.clip)
    assert_close(b16.obs_norm_buf.float(), want, rtol=2**-8, atol=1e-6, what="bf16 normalised obs vs oracle (half a bf16 ulp)")


# ---------------------------------------------------------------------------------------
# the whole env against a recording of the reference's own HumanoidPHC.step / reset
# ---------------------------------------------------------------------------------------
class _ShimUnderReplay:
    """Adapts the CUDA shim to conftest.replay_env_rollout (default config of the reference: power reward,
    frozen hands / toes, reference-state init, AMP observations on)."""

    def __init__(self, g):
        from humanoid_b200 import HumanoidPHC

        clock = clock_from_golden(g)
        N = clock.progress_buf.shape[0]
        self.env = env = HumanoidPHC(
            MotionLib(lib_from_golden(g), device=DEV), N, device=DEV, use_power_reward=True,
            rew_power_coef=float(g.inp("rew_power_coef")), termination_distance=float(g.inp("termination_distance")),
            use_amp_obs=True, num_amp_obs_steps=int(g.inp("num_amp_obs_steps")),
        )  # fmt: skip
        env.set_clock(clock.to(DEV))
        env._pd_action_offset.copy_(g.inp("pd_action_offset"))
        env._pd_action_scale.copy_(g.inp("pd_action_scale"))
        assert torch.equal(env.dof_subset.cpu(), g.inp("dof_subset")) and torch.equal(env._key_body_ids.cpu(), g.inp("key_body_ids"))

    def __getattr__(self, name):
        return getattr(self.env, name)

    def write_sim(self, state, dof_state, dof_force):
        env = self.env
        env._rigid_body_state_reshaped.copy_(state)
        env._humanoid_root_states.copy_(state[:, 0])
        env._dof_state.view(-1, 69, 2).copy_(dof_state)
        env.dof_force_tensor.copy_(dof_force)

    def read_sim(self):
        return self.env._rigid_body_state_reshaped, self.env._humanoid_root_states, self.env._dof_state.view(-1, 69, 2)

    def step(self, actions, physics):
        pd = self.env._action_to_pd_targets(cuda(actions), freeze_hand=True, freeze_toe=True)
        physics(self)
        self.env.step(cuda(actions))
        return pd

    def reset(self, env_ids, phase):
        if len(env_ids):
            self.env.reset(cuda(env_ids), cuda(phase))


def test_env_rollout_vs_reference_step_and_reset(golden):
    from conftest import replay_env_rollout

    g = golden("env_rollout")
    replay_env_rollout(g, _ShimUnderReplay(g), read=lambda t: t.cpu(), tol=OBS_TOL, dof_tol=DOF_DERIVED_TOL)


class _ShimUnderResetModes(_ShimUnderReplay):
    """The shim in StateInit.Default / Hybrid (AMP off), behind conftest.replay_env_reset_modes.  ``route``: the reset
    of the recorded indices through ``reset(env_ids)``, through ``reset_done()`` (device-side mask = reset_buf), or
    inside the step (``auto_reset``: the reset rides behind the step launch, the replay's reset() is a no-op)."""

    def __init__(self, g, mode, route):
        from humanoid_b200 import HumanoidPHC

        clock = clock_from_golden(g)
        N = clock.progress_buf.shape[0]
        self.g, self.mode, self.route, self.k = g, mode, route, 0
        self.env = env = HumanoidPHC(
            MotionLib(lib_from_golden(g), device=DEV), N, device=DEV, use_power_reward=True,
            rew_power_coef=float(g.inp("rew_power_coef")), termination_distance=float(g.inp("termination_distance")),
        )  # fmt: skip
        env.set_clock(clock.to(DEV))
        env._pd_action_offset.copy_(g.inp("pd_action_offset"))
        env._pd_action_scale.copy_(g.inp("pd_action_scale"))
        env.state_init = mode.lower()
        env.hybrid_init_prob = float(g.inp("hybrid_init_prob"))
        env._initial_humanoid_root_states.copy_(g.inp("initial_root_states"))
        env._initial_dof_pos.copy_(g.inp("initial_dof_pos"))
        env._initial_dof_vel.copy_(g.inp("initial_dof_vel"))
        if route == "auto_reset":
            env.enable_auto_reset(True)

    def _by_env(self, k):
        """The recorded draws of reset k, scattered to [N] (what the device-side routes take)."""
        g, mode, N = self.g, self.mode, self.env.num_envs
        rows = g.inp(f"{mode}.reset_indices.{k}")
        dmask, phase = torch.zeros(N, dtype=torch.bool), torch.zeros(N)
        if mode == "Default":
            dmask[rows] = True
        else:
            ref = g.inp(f"{mode}.ref_mask.{k}")
            dmask[rows] = ~ref
            phase[rows[ref]] = g.inp(f"{mode}.phase.{k}")
        return cuda(phase), cuda(dmask)

    def step(self, actions, physics):
        pd = self.env._action_to_pd_targets(cuda(actions), freeze_hand=True, freeze_toe=True)
        physics(self)
        if self.route == "auto_reset":
            phase, dmask = self._by_env(self.k)
            self.env.step(cuda(actions), phase_by_env=phase, default_mask_by_env=dmask)
            self.after = {n: getattr(self.env, n).clone() for n in ("obs_buf", "progress_buf", "reset_buf", "_terminate_buf")}
            # the replay compares the step's own outputs first: the flags live in extras, obs / progress of the envs that
            # were reset in the same launch are checked after the replay's reset()
            rows = cuda(self.g.inp(f"{self.mode}.reset_indices.{self.k}"))
            assert torch.equal(torch.nonzero(self.env.extras["reset"]).squeeze(-1), rows)
        else:
            self.env.step(cuda(actions))
        return pd

    def reset(self, env_ids, phase, default_mask):
        if self.route == "reset":
            self.env.reset(cuda(env_ids), cuda(phase), cuda(default_mask) if self.mode == "Hybrid" else None)
        elif self.route == "reset_done":
            ph, dm = self._by_env(self.k)
            self.env.reset_done(ph, dm)
        self.k += 1


@pytest.mark.parametrize("route", ["reset", "reset_done", "auto_reset"])
@pytest.mark.parametrize("mode", ["Default", "Hybrid"])
def test_env_reset_modes_vs_reference(golden, mode, route):
    """StateInit.Default / StateInit.Hybrid (envs/humanoid_phc.py:678-745) against a recording of the reference's own
    HumanoidPHC.step / reset in those modes, through all three reset routes of the shim."""
    from conftest import replay_env_reset_modes

    g = golden("env_reset_modes")
    shim = _ShimUnderResetModes(g, mode, route)
    if route != "auto_reset":
        assert replay_env_reset_modes(g, shim, mode, read=lambda t: t.cpu(), tol=OBS_TOL, dof_tol=DOF_DERIVED_TOL) == 3
        return
    # auto_reset: step + reset are one call, so the per-step checks of the replay cannot see the pre-reset rows of the
    # envs that were reset; compare the untouched envs with the step record and everything with the reset record
    K = 3
    for k in range(K):
        def physics(e, k=k):
            e.write_sim(g.inp(f"{mode}.state.{k}"), g.inp(f"{mode}.dof_state.{k}"), g.inp(f"{mode}.dof_force.{k}"))

        shim.step(g.inp(f"{mode}.actions.{k}"), physics)
        env = shim.env
        rows = g.inp(f"{mode}.reset_indices.{k}")
        keep = torch.ones(env.num_envs, dtype=torch.bool)
        keep[rows] = False
        s = lambda n, k=k: g.out(f"{mode}.step.{k}.{n}")  # noqa: E731
        r = lambda n, k=k: g.out(f"{mode}.reset.{k}.{n}")  # noqa: E731
        assert_equal_exact(env.extras["reset"].cpu(), s("reset"), f"reset flags[{k}]")
        assert_equal_exact(env.extras["terminate"].cpu(), s("terminate"), f"terminate flags[{k}]")
        assert_close(env.rew_buf.cpu(), s("rew"), what=f"rew[{k}]", **OBS_TOL)
        assert_close(env.obs_buf.cpu()[keep], s("obs")[keep], what=f"obs of the envs not reset[{k}]", **OBS_TOL)
        for name, attr in (("progress", "progress_buf"), ("reset", "reset_buf"), ("terminate", "_terminate_buf"),
                           ("motion_start_times", "_motion_start_times"), ("global_offset", "_global_offset"),
                           ("motion_start_times_offset", "_motion_start_times_offset")):  # fmt: skip
            assert_equal_exact(getattr(env, attr).cpu(), r(name), f"{name} after reset[{k}]")
        sim, root, dof = shim.read_sim()
        assert_close(sim.cpu(), r("rigid_body_state"), what=f"rigid_body_state after reset[{k}]", **OBS_TOL)
        assert_close(root.cpu()[rows], r("root_states")[rows], what=f"root_states after reset[{k}]", **OBS_TOL)
        assert_close(dof.cpu()[rows], r("dof_state")[rows], what=f"dof_state after reset[{k}]", **DOF_DERIVED_TOL)
        assert_close(env.obs_buf.cpu(), r("obs"), what=f"obs after reset[{k}]", **OBS_TOL)
        shim.k += 1


# ---------------------------------------------------------------------------------------
# motion-library build (phc_motion_build) against the reference's own load_motions
# ---------------------------------------------------------------------------------------
def _clips_from_golden(g):
    import numpy as np

    nf = g.inp("num_frames").tolist()
    starts = np.concatenate([[0], np.cumsum(nf)])
    clips = {}
    for m in range(len(nf)):
        sl = slice(int(starts[m]), int(starts[m + 1]))
        clip = {"root_trans_offset": g.inp("root_trans_offset")[sl], "pose_aa": g.inp("pose_aa")[sl].numpy().copy(),
                "pose_quat_global": g.inp("pose_quat_global")[sl].numpy(), "beta": np.zeros(16)}  # fmt: skip
        if int(g.inp("fps")[m]) != 30:
            clip["fps"] = int(g.inp("fps")[m])
        clips[f"clip{m}"] = clip
    return clips


def _trees_from_golden(g):
    from humanoid_b200.motion_build import SkeletonTree

    names = [f"b{j}" for j in range(24)]
    return [SkeletonTree(names, g.inp("parent_indices"), lt) for lt in g.inp("local_translation")]


BUILD_EXACT = ("grs", "lrs", "gts", "gvs", "_motion_aa", "_motion_lengths", "_motion_num_frames", "_motion_dt",
               "_motion_fps", "length_starts", "_motion_bodies", "_motion_limb_weights")  # fmt: skip


@pytest.mark.parametrize("variant", ["deterministic", "random_heading", "cropped"])
def test_motion_build_vs_reference_load_motions(golden, variant):
    import random

    import numpy as np

    from humanoid_b200.motion_build import MotionLibSMPL

    g = golden("motion_build")
    lib = MotionLibSMPL(_clips_from_golden(g), device=DEV, is_deterministic=variant == "deterministic",
                        max_length=int(g.inp("crop_max_length")) if variant == "cropped" else -1)  # fmt: skip
    np.random.seed(31)  # the seeds make_golden.py ran the reference's loader under: the host side must draw the
    random.seed(7)      # crop starts and the headings in the reference's order to land on the same clips
    lib.load_motions(_trees_from_golden(g), list(g.inp("gender_betas")), g.inp("limb_weights").numpy(), random_sample=False)
    for k in BUILD_EXACT + ("gavs", "dvs", "grvs", "gravs"):
        got, want = getattr(lib, k).cpu(), g.out(f"{variant}.{k.lstrip('_')}")
        assert got.dtype == want.dtype and got.shape == want.shape, k
        if want.dtype == torch.int64:
            assert_equal_exact(got, want, k)
        elif k in BUILD_EXACT and variant == "deterministic":
            # same roundings in the same order as the reference's mixed fp64/fp32 pipeline: no libm call on
            # these outputs, so they are reproduced to the bit
            assert torch.equal(got, want), f"{k}: {(got != want).sum()} of {got.numel()} differ, max {float((got - want).abs().max())}"
        else:  # sin/cos/acos/atan2 of libdevice vs libm / SLEEF on the way
            assert_close(got, want, what=f"{variant}.{k}", **OBS_TOL)
    if variant != "deterministic":  # the heading went into the caller's pose_aa in place (motion_lib.py:783, :794)
        src = np.concatenate([c["pose_aa"] for c in lib._motion_data_list]).astype(np.float32)
        assert np.array_equal(src, lib._motion_aa.cpu().numpy()) and not np.array_equal(src, g.inp("pose_aa").numpy().astype(np.float32))
    # the built library answers queries like one assembled from the reference's tensors
    ids = torch.arange(lib.num_motions(), device=DEV).repeat(3)
    t = torch.rand(ids.shape[0], device=DEV) * lib._motion_lengths[ids]
    ref = MotionLib({k.lstrip("_"): g.out(f"{variant}.{k.lstrip('_')}") for k in BUILD_EXACT + ("gavs", "dvs")}, device=DEV)
    a, b = lib.get_motion_state(ids, t), ref.get_motion_state(ids, t)
    for k in MOTION_KEYS:
        assert_close(a[k].cpu(), b[k].cpu(), what=f"query {k}", **(DOF_TOL if k == "dof_pos" else OBS_TOL))


def test_motion_build_many_clips_vs_oracle():
    """A few hundred ragged clips in one launch; spot clips (first, last, shortest, longest) against the oracle."""
    import numpy as np
    from scipy.spatial.transform import Rotation as R

    from humanoid_b200.motion_build import build_motion_tensors
    from oracle import build_oracle as B

    rng = np.random.default_rng(3)
    g_parents = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
    M = 300
    nf = rng.integers(2, 120, size=M)
    nf[7], nf[11] = 2, 300
    fps = rng.choice([30, 60, 120], size=M)
    F = int(nf.sum())
    quat = R.random(F * 24, random_state=5).as_quat().reshape(F, 24, 4)
    # smooth in time within a clip: blend towards the previous frame and renormalise
    for f in range(1, F):
        quat[f] = 0.2 * quat[f] + 0.8 * quat[f - 1] * np.sign((quat[f] * quat[f - 1]).sum(-1, keepdims=True))
        quat[f] /= np.linalg.norm(quat[f], axis=-1, keepdims=True)
    trans = np.cumsum(rng.normal(size=(F, 3)) * 0.02, 0)
    aa = rng.normal(size=(F, 72))
    lt = (rng.normal(size=(M, 24, 3)) * 0.2).astype(np.float32)
    u = rng.random(M)
    for heading in (None, u):
        out = build_motion_tensors(quat, trans, aa, nf, fps, g_parents, lt, heading_u=heading, device=DEV)
        torch.cuda.synchronize()
        starts = np.concatenate([[0], np.cumsum(nf)])
        for m in (0, 7, 11, M - 1):
            sl = slice(int(starts[m]), int(starts[m + 1]))
            want = B.build_motion_library(quat[sl], trans[sl], aa[sl], [nf[m]], [fps[m]], g_parents, lt[m : m + 1],
                                          np.zeros((1, 17)), np.zeros((1, 10)),
                                          heading_u=None if heading is None else [u[m]])  # fmt: skip
            for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa"):
                # random i.i.d.-ish rotations: angular velocities reach ~100 rad/s, so a relative bound
                assert_close(out[k][sl].cpu(), want[k], what=f"clip {m} {k}", rtol=2e-5, atol=2e-5)


def test_motion_build_rejects_bad_arguments():
    import numpy as np

    from humanoid_b200 import _cabi
    from humanoid_b200.motion_build import build_motion_tensors

    q = np.zeros((4, 24, 4))
    q[..., 3] = 1
    par = [-1] + list(range(23))
    with pytest.raises(_cabi.PhcError):  # a parent that follows its child
        build_motion_tensors(q, np.zeros((4, 3)), None, [4], [30], [-1, 2, 1] + list(range(2, 23)), np.zeros((1, 24, 3)), device=DEV)
    with pytest.raises(_cabi.PhcError):  # frame count mismatch
        build_motion_tensors(q, np.zeros((4, 3)), None, [5], [30], par, np.zeros((1, 24, 3)), device=DEV)
    with pytest.raises(_cabi.PhcError):
        build_motion_tensors(q, np.zeros((4, 3)), None, [4], [30], par, np.zeros((1, 24, 3)), device="cpu")
    out = build_motion_tensors(q, np.zeros((4, 3)), None, [4], [30], par, np.zeros((1, 24, 3)), device=DEV)
    assert float(out["gvs"].abs().max()) == 0 and float(out["motion_aa"].abs().max()) == 0


def test_resample_motions_rebuilds_the_library_and_resets_vs_oracle(golden):
    """HumanoidPHC.resample_motions (:1363-1379): load_motions on the device, xy re-anchoring of the global
    offset, reset of every env — against the oracle env on the oracle-built library, with a step either side."""
    import numpy as np

    from humanoid_b200 import HumanoidPHC
    from humanoid_b200.motion_build import MotionLibSMPL
    from oracle import build_oracle as B

    g = golden("motion_build")
    clips, trees = _clips_from_golden(g), _trees_from_golden(g)
    M = len(clips)
    N = 2 * M  # env e holds clip (e + start_idx) % M and env e's skeleton
    trees = [trees[e % M] for e in range(N)]
    shapes, limbs = g.inp("gender_betas").repeat(2, 1), g.inp("limb_weights").repeat(2, 1).float()
    starts = np.concatenate([[0], np.cumsum(g.inp("num_frames").tolist())])

    def oracle_lib(order):
        sl = [slice(int(starts[m]), int(starts[m + 1])) for m in order]
        cat = lambda k: np.concatenate([g.inp(k).numpy()[s] for s in sl])  # noqa: E731
        return O.OracleMotionLib(B.build_motion_library(
            cat("pose_quat_global"), cat("root_trans_offset"), cat("pose_aa"), [int(g.inp("num_frames")[m]) for m in order],
            [int(g.inp("fps")[m]) for m in order], g.inp("parent_indices").tolist(),
            np.stack([t.local_translation.numpy() for t in trees]), shapes.numpy(), limbs.numpy()))  # fmt: skip

    lib = MotionLibSMPL(clips, device=DEV, is_deterministic=True)
    lib.load_motions(trees, list(shapes), limbs.numpy(), start_idx=0)
    env = HumanoidPHC(lib, N, device=DEV, use_amp_obs=True)
    env.set_humanoid_assets(trees, shapes, limbs)
    order0 = [e % M for e in range(N)]
    zeros = torch.zeros(N)
    ref = O.OracleEnv(oracle_lib(order0), N, torch.zeros(N, dtype=torch.short), zeros, zeros, torch.zeros(N, 3),
                      torch.arange(N), torch.zeros(69), torch.ones(69), env.dof_subset.cpu(), env._key_body_ids.cpu(),
                      rew_power_coef=0.0)  # fmt: skip
    gen = torch.Generator().manual_seed(4)
    phase = torch.rand(N, generator=gen)
    env.reset(phase=cuda(phase))
    ref.reset(torch.arange(N), phase)

    def physics(e):  # a drifted copy of the posed state stands in for PhysX
        e.state[:] = e.state + 0.01
        e.root_states[:] = e.state[:, 0]

    def step_both():
        ref.step(torch.zeros(N, 69), physics)
        env._rigid_body_state_reshaped.copy_(ref.state)
        env._humanoid_root_states.copy_(ref.root_states)
        env._dof_state.view(N, 69, 2).copy_(ref.dof_state)
        env.step()
        assert_equal_exact(env.reset_buf.cpu(), ref.reset_buf, "reset_buf")
        assert_close(env.obs_buf.cpu(), ref.obs_buf, what="obs", **OBS_TOL)

    step_both()
    # the loader's start_idx is host state of the caller in the reference too (forward_motion_samples :1391-1402);
    # resample with the sequential branch so both sides load the same clips
    phase2 = torch.rand(N, generator=gen)
    env.resample_motions(seq_motions=True, phase=cuda(phase2))
    ref.resample_motions(oracle_lib(order0), phase2)
    assert_close(env._global_offset.cpu(), ref._global_offset, what="global offset", rtol=1e-6, atol=1e-6)
    assert_close(env._motion_start_times.cpu(), ref._motion_start_times, what="start times", rtol=0, atol=0)
    assert_close(env._rigid_body_state_reshaped.cpu(), ref.state, what="posed state", **OBS_TOL)
    assert_close(env.obs_buf.cpu(), ref.obs_buf, what="obs after resample", **OBS_TOL)
    assert_close(env.amp_obs.cpu().view(N, -1), ref._amp_obs_buf.view(N, -1), what="amp obs", **DOF_DERIVED_TOL)
    step_both()


def test_eval_sweep_swaps_libraries_and_walks_the_clips(golden):
    """toggle_eval_mode -> begin_seq_motion_samples -> forward_motion_samples -> untoggle_eval_mode
    (humanoid_phc.py:1381-1455) with a train and an eval library (longest clip first) over the same clips."""
    import numpy as np

    from humanoid_b200 import HumanoidPHC
    from humanoid_b200.motion_build import MotionLibSMPL
    from oracle import build_oracle as B

    g = golden("motion_build")
    M, N = 6, 4
    trees = _trees_from_golden(g)[:N]
    shapes, limbs = g.inp("gender_betas")[:N], g.inp("limb_weights")[:N].float()
    nf = g.inp("num_frames").tolist()
    starts = np.concatenate([[0], np.cumsum(nf)])
    eval_order = sorted(range(M), key=lambda m: -nf[m])  # stable: load_data's im_eval sort (motion_lib.py:209-217)

    def oracle_lib(order):
        sl = [slice(int(starts[m]), int(starts[m + 1])) for m in order]
        cat = lambda k: np.concatenate([g.inp(k).numpy()[s] for s in sl])  # noqa: E731
        return O.OracleMotionLib(B.build_motion_library(
            cat("pose_quat_global"), cat("root_trans_offset"), cat("pose_aa"), [nf[m] for m in order],
            [int(g.inp("fps")[m]) for m in order], g.inp("parent_indices").tolist(),
            np.stack([t.local_translation.numpy() for t in trees]), shapes.numpy(), limbs.numpy()))  # fmt: skip

    train = MotionLibSMPL(_clips_from_golden(g), device=DEV, is_deterministic=True)
    evall = MotionLibSMPL(_clips_from_golden(g), device=DEV, im_eval=True)
    train.load_motions(trees, list(shapes), limbs.numpy(), random_sample=False)
    env = HumanoidPHC(train, N, device=DEV)
    env.set_humanoid_assets(trees, shapes, limbs)
    env.set_motion_libs(train, evall)
    assert evall._motion_data_keys.tolist() == [f"clip{m}" for m in eval_order]

    gen = torch.Generator().manual_seed(11)
    for sweep, start in enumerate((0, N)):
        phase = torch.rand(N, generator=gen)
        if sweep == 0:
            assert env.toggle_eval_mode(phase=cuda(phase)) == M and env._motion_lib is evall
        else:
            env.resample_motions(phase=cuda(phase))  # flag_test: forward_motion_samples (:1364-1365)
        assert env.motion_sample_start_idx == start
        order = [eval_order[(e + start) % M] for e in range(N)]
        assert env.current_motion_ids.tolist() == [(e + start) % M for e in range(N)]
        assert env.get_motion_steps().tolist() == [int(np.ceil(nf[m] * 30 / int(g.inp("fps")[m]))) for m in order]
        zeros = torch.zeros(N)
        ref = O.OracleEnv(oracle_lib(order), N, torch.zeros(N, dtype=torch.short), zeros, zeros, torch.zeros(N, 3),
                          torch.arange(N), torch.zeros(69), torch.ones(69), env.dof_subset.cpu(), env._key_body_ids.cpu(),
                          use_amp_obs=False)  # fmt: skip
        ref.flag_test = True
        ref.reset(torch.arange(N), phase)
        assert_close(env._motion_start_times.cpu(), ref._motion_start_times, what="start times", rtol=0, atol=0)
        assert_close(env._rigid_body_state_reshaped.cpu(), ref.state, what="posed state", **OBS_TOL)
        assert_close(env.obs_buf.cpu(), ref.obs_buf, what="obs", **OBS_TOL)
    hist = env.untoggle_eval_mode(failed_keys=["clip1", "clip4"])
    assert env._motion_lib is train and not env.flag_test
    assert hist.tolist() == [0, 1, 0, 0, 1, 0] and train._sampling_prob.tolist() == [0, 0.5, 0, 0, 0.5, 0]
    assert float(env._termination_distances[0]) == 0.25


def test_motion_build_large_ragged_library_properties():
    """Size-independent properties at ~0.5 M frames (3000 ragged clips, mixed fps), where the oracle would take minutes:
    rigid translation at constant velocity with constant joint rotations must give exact kinematics."""
    import numpy as np
    from scipy.spatial.transform import Rotation as R

    from humanoid_b200.motion_build import build_motion_tensors

    rng = np.random.default_rng(11)
    parents = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
    M = 3000
    nf = rng.integers(2, 340, size=M)
    nf[:3] = (2, 3, 7000)  # shortest clips and one AMASS-length clip
    fps = rng.choice([30, 60, 120], size=M)
    F = int(nf.sum())
    starts = np.concatenate([[0], np.cumsum(nf)])
    clip = np.repeat(np.arange(M), nf)
    k = np.arange(F) - starts[clip]  # frame index inside the clip
    vel = rng.normal(size=(M, 3)) * 0.01  # positions stay within a few metres even over the 7000-frame clip
    x0 = rng.normal(size=(M, 3))
    trans = x0[clip] + vel[clip] * (k / fps[clip])[:, None]
    pose = R.random(M * 24, random_state=2).as_quat().reshape(M, 24, 4)  # one constant pose per clip
    pose *= np.where(pose[..., 3:] < 0, -1.0, 1.0)  # w >= 0, the loader's canonical sign (quat_pos)
    quat = pose[clip]
    lt = (rng.normal(size=(M, 24, 3)) * 0.2).astype(np.float32)
    out = build_motion_tensors(quat, trans, None, nf, fps, parents, lt, device=DEV)
    torch.cuda.synchronize()
    assert out["gts"].shape == (F, 24, 3) and torch.equal(out["length_starts"].cpu(), torch.from_numpy(starts[:-1]))
    assert torch.equal(out["grs"].cpu(), torch.from_numpy(quat).float()), "global rotations are the inputs, rounded"
    assert torch.equal(out["gts"][:, 0].cpu(), torch.from_numpy(trans).float()), "root = root translation, rounded"
    # constant pose: no angular or dof velocity anywhere, bodies ride along with the root
    assert float(out["gavs"].abs().max()) == 0.0 and float(out["dvs"].abs().max()) < 1e-3
    rel = (out["gts"] - out["gts"][:, :1]).cpu()
    first = torch.from_numpy(starts[:-1])
    assert float((rel - rel[first][clip]).abs().max()) < 2e-5, "body offsets from the root are constant inside a clip"
    # constant root velocity survives np.gradient and the 17-tap filter with 'nearest' edges (taps sum to 1)
    want = torch.from_numpy(vel[clip]).float()[:, None, :].expand(-1, 24, -1)
    err = (out["gvs"].cpu() - want).abs()
    tol = 2e-5 + fps[clip][:, None, None] * float(np.spacing(np.float32(np.abs(trans).max() + 1.0)))  # one fp32 position ulp times fps
    assert bool((err <= torch.from_numpy(tol)).all()), float(err.max())
    # local rotations compose back to the global ones through the tree
    lrs, grs = out["lrs"].cpu().double(), out["grs"].cpu().double()
    j = 7
    comp = O.quat_mul(grs[:, parents[j]], lrs[:, j])
    sign = torch.sign((comp * grs[:, j]).sum(-1, keepdim=True))
    assert float((comp * sign - grs[:, j]).abs().max()) < 1e-6


# ---------------------------------------------------------------------------------------
# no kernel writes outside its outputs (compute-sanitizer is not available on the pool): every output buffer of
# the env is re-seated in the middle of a canary-filled allocation
# ---------------------------------------------------------------------------------------
def _guarded(t, pad=4096):
    """A tensor like ``t`` carved out of a larger allocation whose bytes before and after it hold a canary."""
    nbytes = t.numel() * t.element_size()
    lead = pad + (-pad) % 256
    raw = torch.full((lead + nbytes + pad,), 0xA5, dtype=torch.uint8, device=t.device)
    view = raw[lead : lead + nbytes].view(t.dtype).view(t.shape)
    view.copy_(t)
    return raw, view, lead, nbytes


@pytest.mark.parametrize("N,T,mode", [(4099, 1, "fast"), (4096, 1, "fast_norm16"), (4099, 1, "fast_norm32"), (1001, 1, "generic"),
                                      (515, 3, "multi"), (515, 3, "generic_norm")])  # fmt: skip
def test_step_kernels_stay_inside_their_output_buffers(N, T, mode):
    from humanoid_b200 import HumanoidPHC, RunningNorm, _cabi

    lib_data, clock, state = _gpu_case(N, 48, 321, max_progress=30)
    env = HumanoidPHC(MotionLib(lib_data, device=DEV), N, device=DEV, time_steps=T, obs_moments=True)
    env.set_sim_state(state)
    env.set_clock(clock)
    if "norm" in mode:
        env.set_obs_normalizer(RunningNorm(env.num_obs, device=DEV), dtype=torch.bfloat16 if "16" in mode else torch.float32)
    guards = {}
    for name in ("obs_buf", "rew_buf", "reward_raw", "reset_buf", "_terminate_buf", "progress_buf", "obs_norm_buf",
                 "_obs_moment_buckets"):  # fmt: skip
        t = getattr(env, name)
        if t is None:
            continue
        raw, view, lead, nbytes = _guarded(t)
        setattr(env, name, view)
        guards[name] = (raw, lead, nbytes)
    env._step_args = None
    capi = _cabi.load()
    try:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 1 if mode.startswith("generic") else 0)
        for _ in range(3):
            env.step()
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
    torch.cuda.synchronize()
    assert float(env.obs_buf.abs().sum()) > 0 and int(env.progress_buf.max()) >= 3
    for name, (raw, lead, nbytes) in guards.items():
        assert bool((raw[:lead] == 0xA5).all()), f"{name}: bytes before the buffer were written"
        assert bool((raw[lead + nbytes :] == 0xA5).all()), f"{name}: bytes after the buffer were written"


def test_build_and_reset_kernels_stay_inside_their_output_buffers(golden):
    """Same for the per-function outputs that have ragged shapes: the motion-library build (dvs rows are 23 joints,
    motion_aa 72) and the device-side reset of a few envs."""
    import ctypes as C

    import numpy as np

    from humanoid_b200 import HumanoidPHC, _cabi
    from humanoid_b200.motion_build import gaussian_taps

    g = golden("motion_build")
    F, M = int(g.inp("num_frames").sum()), len(g.inp("num_frames"))
    starts = np.concatenate([[0], np.cumsum(g.inp("num_frames").tolist())[:-1]]).astype(np.int64)
    dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(DEV)  # noqa: E731
    ins = [dev(g.inp("pose_quat_global").numpy(), torch.float64), dev(g.inp("root_trans_offset").numpy(), torch.float64),
           dev(g.inp("pose_aa").numpy(), torch.float64), dev(g.inp("local_translation").numpy(), torch.float32),
           dev(g.inp("num_frames").numpy(), torch.int64), dev(starts, torch.int64), dev(g.inp("fps").numpy(), torch.float64)]  # fmt: skip
    shapes = {"gts": (F, 24, 3), "grs": (F, 24, 4), "lrs": (F, 24, 4), "gvs": (F, 24, 3), "gavs": (F, 24, 3), "dvs": (F, 23, 3),
              "motion_aa": (F, 72)}  # fmt: skip
    outs = {k: _guarded(torch.zeros(v, device=DEV)) for k, v in shapes.items()}
    scratch = _guarded(torch.zeros(F * 108, dtype=torch.float64, device=DEV))
    parents = (C.c_int32 * 24)(*g.inp("parent_indices").tolist())
    taps = (C.c_double * 17)(*gaussian_taps().tolist())
    args = _cabi.PhcBuildArgs(*[t.data_ptr() for t in ins], None, parents, taps, F, M,
                              *[outs[k][1].data_ptr() for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa")],
                              scratch[1].data_ptr())  # fmt: skip
    _cabi.check(_cabi.load().phc_motion_build(C.byref(args), torch.cuda.current_stream().cuda_stream), "phc_motion_build")
    torch.cuda.synchronize()
    assert_close(outs["gts"][1].cpu(), g.out("deterministic.gts"), rtol=0, atol=0, what="gts")
    for name, (raw, _, lead, nbytes) in list(outs.items()) + [("scratch", scratch)]:
        assert bool((raw[:lead] == 0xA5).all()) and bool((raw[lead + nbytes :] == 0xA5).all()), name

    lib_data, clock, state = _gpu_case(37, 5, 77, max_progress=10)
    env = HumanoidPHC(MotionLib(lib_data, device=DEV), 37, device=DEV, use_amp_obs=True)
    env.set_clock(clock)
    guards = {}
    for name in ("obs_buf", "progress_buf", "reset_buf", "_terminate_buf", "_motion_start_times", "_global_offset",
                 "_humanoid_root_states", "_rigid_body_state_reshaped", "_amp_obs_buf", "_amp_obs_demo_buf"):  # fmt: skip
        raw, view, lead, nbytes = _guarded(getattr(env, name))
        setattr(env, name, view)
        guards[name] = (raw, lead, nbytes)
    env._bind_body_views()
    env._curr_amp_obs_buf, env._hist_amp_obs_buf = env._amp_obs_buf[:, 0], env._amp_obs_buf[:, 1:]
    env._step_args = None
    env.reset(torch.tensor([0, 5, 36], device=DEV), torch.tensor([0.1, 0.5, 0.9], device=DEV))
    env.step()
    torch.cuda.synchronize()
    for name, (raw, lead, nbytes) in guards.items():
        assert bool((raw[:lead] == 0xA5).all()) and bool((raw[lead + nbytes :] == 0xA5).all()), name


# ---------------------------------------------------------------------------------------
# randomised differential test: many small seeded workloads with adversarial clocks against the oracle
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(16))
def test_fused_step_randomised_differential(seed):
    """Random batch size / clip count / fps mix / id regime / time alignment / rotation regime per seed, clocks that run
    past the end of short clips (pass_time) and start before 0, T in {1, 2, 5}, every step kernel (TMA kernels by
    default, the generic kernel for odd seeds): flags and progress exact, floats within the 1e-5 bar, two steps deep."""
    from humanoid_b200 import _cabi

    rng = torch.Generator().manual_seed(1000 + seed)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))  # noqa: E731
    N, M = r(1, 333), r(1, 24)
    T = (1, 1, 2, 5)[seed % 4]
    kw = dict(num_envs=N, num_motions=M, seed=2000 + seed, min_frames=r(2, 12), max_frames=r(13, 80),
              fps_choices=((30,), (30, 60), (30, 60, 120))[seed % 3], ids=("mod", "random")[seed % 2],
              aligned=bool(seed % 3), rot_regime=("mocap", "random")[(seed // 2) % 2], max_progress=r(0, 60))  # fmt: skip
    lib_data, clock, state = make_case_cpu(**kw)
    # adversarial clocks: some envs start before the clip (negative time), some far past its end
    n_adv = max(1, N // 7)
    idx = torch.randperm(N, generator=rng)[:n_adv]
    clock.motion_start_times_offset[idx[: n_adv // 2 + 1]] = -1.5
    clock.motion_start_times[idx[n_adv // 2 :]] += 40.0
    capi = _cabi.load()
    env = env_from(lib_data, clock, state, time_steps=T)
    prog = clock.progress_buf.clone()
    lib_o = O.OracleMotionLib(lib_data)
    try:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, seed % 2)
        for step in range(2):
            want = O.step(lib_o, state, prog, clock.motion_start_times, clock.motion_start_times_offset, clock.global_offset,
                          clock.sampled_motion_ids, torch.full((24,), 0.25), synth.SIM_DT, time_steps=T)  # fmt: skip
            env.step()
            what = f"seed {seed} step {step} {kw} T={T}"
            assert_equal_exact(env.progress_buf, prog, what + ": progress")
            assert_equal_exact(env.reset_buf, want[3], what + ": reset")
            assert_equal_exact(env._terminate_buf, want[4], what + ": terminated")
            assert_close(env.obs_buf, want[0], what=what + ": obs", **OBS_TOL)
            assert_close(env.rew_buf, want[1], what=what + ": reward", **OBS_TOL)
            assert_close(env.reward_raw[:, :4], want[2], what=what + ": reward_raw", **OBS_TOL)
    finally:
        capi.phc_set_option(_cabi.OPT_FORCE_GENERIC_STEP, 0)
