"""Generate the golden fixtures in this directory by running the REFERENCE's own functions.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is an ``.npz`` holding seeded inputs (``in.*``) and the outputs
(``out.*``) of the unmodified reference functions on CPU fp32:
``MotionLibBase.get_motion_state`` / ``_calc_frame_blend`` (PHC/motion_lib.py:549,655),
``compute_humanoid_observations_smpl_max`` / ``compute_imitation_observations_v6`` /
``compute_imitation_reward`` / ``compute_humanoid_im_reset`` (PHC/envs/common.py:23,107,
271,326), ``torch_utils.slerp`` (:110) and ``RunningNorm.update``
(PHC/policies/running_norm.py:23), orchestrated as ``HumanoidPHC.step`` does
(PHC/envs/humanoid_phc.py:138-149).  The reference has no tests or vectors of its own
(SURVEY §4), so these are the pin for ``oracle/phc_oracle.py`` and for the CUDA path.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from humanoid_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

ref_tu, ref_common, ref_ml = ref_loader.load()

RWD = dict(
    k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1,
    imitation_reward_dim=4, full_body_reward=True, use_power_reward=True,
)  # fmt: skip
DT = synth.SIM_DT
EVAL_BODY_IDS = [i for i in range(24) if i not in (4, 8, 18, 23)]  # body_sets.py:57 (no hands/toes)

MOTION_KEYS = (
    "root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa",
    "rg_pos", "rb_rot", "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights",
)  # fmt: skip


def npify(d, prefix):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[f"{prefix}.{k}"] = v.detach().cpu().numpy()
        else:
            out[f"{prefix}.{k}"] = np.asarray(v)
    return out


def ref_query(lib_data, ids, times, offset):
    lib = ref_loader.make_reference_lib(lib_data)
    return lib.get_motion_state(ids, times, offset)


def reference_step(lib, state, clock, term_dist, time_steps=1, reset_body_ids=None, use_mean=False, early=True):
    """The post-physics half of HumanoidPHC.step with the reference's functions."""
    J = 24
    pos, rot, vel, ang = synth.body_views(state, J)
    progress = clock.progress_buf.clone()
    progress += 1  # humanoid_phc.py:138
    out = {}

    t = progress * DT + clock.motion_start_times + clock.motion_start_times_offset  # :1236
    ref = lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    motion_len = lib._motion_lengths[clock.sampled_motion_ids]
    nf = lib._motion_num_frames[clock.sampled_motion_ids]
    mdt = lib._motion_dt[clock.sampled_motion_ids]
    i0, i1, bl = lib._calc_frame_blend(t, motion_len, nf, mdt)
    out.update({"t0": t, "t0.frame_idx0": i0, "t0.frame_idx1": i1, "t0.blend": bl})
    for k in MOTION_KEYS:
        out[f"t0.{k}"] = ref[k]

    reward, raw = ref_common.compute_imitation_reward(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang,
        ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], RWD,
    )  # fmt: skip
    out["reward"], out["reward_raw"] = reward, raw

    rb = torch.arange(J) if reset_body_ids is None else torch.tensor(reset_body_ids)
    pass_time = t >= motion_len  # :1317 (ids == arange there; length[id] in general)
    reset, term = ref_common.compute_humanoid_im_reset(
        torch.ones(state.shape[0], dtype=torch.bool), progress,
        torch.zeros(state.shape[0], J, 3), torch.zeros(4, dtype=torch.long),
        pos[:, rb].clone(), ref["rg_pos"][:, rb].clone(), pass_time, early, term_dist[rb], use_mean,
    )  # fmt: skip
    out["reset"], out["terminated"], out["pass_time"] = reset, term, pass_time

    self_obs = ref_common.compute_humanoid_observations_smpl_max(
        pos, rot, vel, ang, None, None, True, True, True, False, False
    )
    out["self_obs"] = self_obs

    refs = []
    for k in range(1, time_steps + 1):
        tk = (progress + k) * DT + clock.motion_start_times + clock.motion_start_times_offset  # :1063-1067
        r = lib.get_motion_state(clock.sampled_motion_ids, tk, clock.global_offset)
        refs.append(r)
        if k == 1:
            i0, i1, bl = lib._calc_frame_blend(tk, motion_len, nf, mdt)
            out.update({"t1": tk, "t1.frame_idx0": i0, "t1.frame_idx1": i1, "t1.blend": bl})
            for key in MOTION_KEYS:
                out[f"t1.{key}"] = r[key]

    def stack(key):
        return torch.stack([r[key] for r in refs], 1).reshape((-1,) + refs[0][key].shape[1:])

    task = ref_common.compute_imitation_observations_v6(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang,
        stack("rg_pos"), stack("rb_rot"), stack("body_vel"), stack("body_ang_vel"), time_steps, True,
    )  # fmt: skip
    out["task_obs"] = task
    out["obs"] = torch.cat([self_obs, task], dim=-1)  # :949
    out["progress_after"] = progress
    return out


def case_reset_mean():
    """compute_humanoid_im_reset(use_mean=True) (PHC/envs/common.py:343-346) on rows whose mean distance sits within
    a few ulps of the threshold: the flag then depends on the ASSOCIATION of ATen's row sum, which is what the CUDA
    kernels and oracle/phc_oracle_int.c have to reproduce.  One block of rows per reset-body count R."""
    g = torch.Generator().manual_seed(21)
    thr = 0.15
    arrays = {"in.threshold": np.float32(thr), "in.rs": np.asarray([1, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17, 20, 23, 24])}
    for R in arrays["in.rs"].tolist():
        n = 192
        ref = torch.randn(n, R, 3, generator=g)
        delta = torch.randn(n, R, 3, generator=g) * 0.1
        for _ in range(3):  # rescale so that the mean distance lands on the threshold to within rounding
            m = torch.norm((ref + delta) - ref, dim=-1).mean(dim=-1)
            delta = delta * (thr / m)[:, None, None]
        # spread the rows over the few floats either side of the threshold
        delta = delta * (1.0 + torch.randint(-2, 4, (n,), generator=g).float() * 6e-8)[:, None, None]
        pos = ref + delta
        progress = torch.full((n,), 5, dtype=torch.int16)
        progress[::17] = 1  # has_fallen *= progress_buf > 1
        pass_time = torch.zeros(n, dtype=torch.bool)
        pass_time[::13] = True
        term_dist = torch.full((R,), thr, dtype=torch.float32)
        reset, term = ref_common.compute_humanoid_im_reset(
            torch.ones(n, dtype=torch.bool), progress, torch.zeros(n, R, 3), torch.zeros(4, dtype=torch.long),
            pos.clone(), ref.clone(), pass_time, True, term_dist, True,
        )  # fmt: skip
        mean = torch.norm(pos - ref, dim=-1).mean(dim=-1)
        frac_on = float(((mean - thr).abs() <= 3e-8).float().mean())
        print(f"  R={R}: terminated {float(term.float().mean()):.2f}, within 2 ulp of the threshold {frac_on:.2f}")
        arrays.update({f"in.pos{R}": pos.numpy(), f"in.ref{R}": ref.numpy(), f"in.progress{R}": progress.numpy(),
                       f"in.pass_time{R}": pass_time.numpy(), f"out.reset{R}": reset.numpy(),
                       f"out.terminated{R}": term.numpy(), f"out.mean{R}": mean.numpy()})  # fmt: skip
    save("reset_mean", arrays)


def save(name, arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


def case_step(name, N, M, seed, time_steps=1, reset_body_ids=None, use_mean=False, early=True, term=0.25, **kw):
    lib_data, clock, state = synth.make_case(N, M, ref_query, seed=seed, device="cpu", **kw)
    lib = ref_loader.make_reference_lib(lib_data)
    term_dist = torch.full((24,), term, dtype=torch.float32)
    out = reference_step(lib, state, clock, term_dist, time_steps, reset_body_ids, use_mean, early)
    arrays = {}
    arrays.update(npify(lib_data.as_dict(), "in.lib"))
    arrays.update(npify(clock.__dict__, "in.clock"))
    arrays["in.state"] = state.numpy()
    arrays["in.term_dist"] = term_dist.numpy()
    arrays["in.time_steps"] = np.asarray(time_steps)
    arrays["in.use_mean"] = np.asarray(use_mean)
    arrays["in.early"] = np.asarray(early)
    arrays["in.reset_body_ids"] = np.asarray(list(range(24)) if reset_body_ids is None else reset_body_ids)
    arrays.update(npify(out, "out"))
    save(name, arrays)


def case_frame_blend():
    """Adversarial times for the integer path: exact frame boundaries and +-1 ulp around
    them, negatives, beyond the clip, zero; 30/60/120 fps; 2-frame and long clips."""
    rows = []
    for fps in (30.0, 60.0, 120.0):
        dt = 1.0 / fps
        for nf in (2, 3, 17, 60, 300, 7000):
            length = dt * (nf - 1)
            ks = sorted(set([0, 1, 2, nf // 2, nf - 2, nf - 1, nf, nf + 5]))
            for k in ks:
                base = np.float32(k * (1 / 30)) if fps == 30.0 else np.float32(k * dt)
                for t in (base, np.nextafter(base, np.float32(np.inf)), np.nextafter(base, np.float32(-np.inf))):
                    rows.append((float(t), length, nf, dt))
            for t in (-1.0, -1e-8, 0.0, length, length * 1.0000001, length * 2, 1e-9, 0.5 * dt, 1.5 * dt):
                rows.append((t, length, nf, dt))
    g = torch.Generator().manual_seed(5)
    for _ in range(4000):
        fps = (30.0, 60.0, 120.0)[int(torch.randint(0, 3, (1,), generator=g))]
        nf = int(torch.randint(2, 400, (1,), generator=g))
        dt = 1.0 / fps
        length = dt * (nf - 1)
        # progress*dt + start, the env's op order, on frame-aligned starts
        k = int(torch.randint(0, nf, (1,), generator=g))
        p = int(torch.randint(0, 300, (1,), generator=g))
        t = float((torch.tensor([p], dtype=torch.int16) * DT + torch.tensor([k * (1 / 30)], dtype=torch.float32))[0])
        rows.append((t, length, nf, dt))
    time = torch.tensor([r[0] for r in rows], dtype=torch.float32)
    length = torch.tensor([r[1] for r in rows], dtype=torch.float32)
    nf = torch.tensor([r[2] for r in rows], dtype=torch.int64)
    dt = torch.tensor([r[3] for r in rows], dtype=torch.float32)
    lib = object.__new__(ref_ml.MotionLibBase)
    i0, i1, bl = lib._calc_frame_blend(time, length, nf, dt)
    save(
        "frame_blend",
        {
            "in.time": time.numpy(), "in.len": length.numpy(), "in.num_frames": nf.numpy(), "in.dt": dt.numpy(),
            "out.frame_idx0": i0.numpy(), "out.frame_idx1": i1.numpy(), "out.blend": bl.numpy(),
        },
    )  # fmt: skip


def case_slerp():
    g = torch.Generator().manual_seed(9)
    n = 4096
    q0 = torch.randn(n, 4, generator=g)
    q0 = q0 / q0.norm(dim=-1, keepdim=True)
    # small relative rotations of graded size, so every branch is hit
    ang = torch.exp(torch.rand(n, 1, generator=g) * 16 - 16)  # e^-16 .. 1 rad
    axis = torch.randn(n, 3, generator=g)
    axis = axis / axis.norm(dim=-1, keepdim=True)
    dq = torch.cat([axis * torch.sin(ang / 2), torch.cos(ang / 2)], -1)
    q1 = ref_tu.quat_mul(q0, dq)
    q1[::7] = -q1[::7]  # dot < 0
    q1[::11] = q0[::11]  # identical -> |cos| >= 1 or sin ~ 0
    q1[5::97] = q0[5::97] * 1.001  # non-unit, cos > 1 -> acos NaN masked by last select
    rnd = torch.randn(n // 4, 4, generator=g)
    q1[: n // 4] = rnd / rnd.norm(dim=-1, keepdim=True)  # large angles
    t = torch.rand(n, 1, generator=g)
    t[::13] = 0.0
    t[1::13] = 1.0
    out = ref_tu.slerp(q0, q1, t)
    em = ref_tu.quat_to_exp_map(out / out.norm(dim=-1, keepdim=True).clamp_min(1e-9))
    save("slerp", {"in.q0": q0.numpy(), "in.q1": q1.numpy(), "in.t": t.numpy(), "out.q": out.numpy(),
                   "in.qn": (out / out.norm(dim=-1, keepdim=True).clamp_min(1e-9)).numpy(), "out.exp_map": em.numpy()})  # fmt: skip


def case_flags():
    """Flag variants of the standalone functions that the env's default path never takes."""
    lib_data, clock, state = synth.make_case(40, 5, ref_query, seed=99, device="cpu", min_frames=12, max_frames=30)
    lib = ref_loader.make_reference_lib(lib_data)
    pos, rot, vel, ang = synth.body_views(state, 24)
    g = torch.Generator().manual_seed(3)
    smpl = torch.randn(40, 11, generator=g)
    limb = torch.randn(40, 10, generator=g)
    arrays = {"in.state": state.numpy(), "in.smpl": smpl.numpy(), "in.limb": limb.numpy()}
    variants = {
        "default": (True, True, True, False, False),
        "not_upright": (True, True, False, False, False),
        "global_root": (False, True, True, False, False),
        "no_height": (True, False, True, False, False),
        "with_params": (True, True, True, True, True),
        "all_off": (False, False, False, True, False),
    }
    for name, fl in variants.items():
        o = ref_common.compute_humanoid_observations_smpl_max(pos, rot, vel, ang, smpl, limb, *fl)
        arrays[f"out.self.{name}"] = o.numpy()
        arrays[f"in.self.{name}"] = np.asarray(fl)
    t = synth.reward_time(clock, extra_steps=1)
    ref = lib.get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel"):
        arrays[f"in.ref.{k}"] = ref[k].numpy()
    # v6, not upright; and a 12-body tracking subset
    o = ref_common.compute_imitation_observations_v6(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], 1, False
    )  # fmt: skip
    arrays["out.v6.not_upright"] = o.numpy()
    sub = torch.arange(0, 24, 2)
    arrays["in.subset"] = sub.numpy()
    o = ref_common.compute_imitation_observations_v6(
        pos[:, 0], rot[:, 0], pos[:, sub], rot[:, sub], vel[:, sub], ang[:, sub],
        ref["rg_pos"][:, sub], ref["rb_rot"][:, sub], ref["body_vel"][:, sub], ref["body_ang_vel"][:, sub], 1, True,
    )  # fmt: skip
    arrays["out.v6.subset12"] = o.numpy()
    r, raw = ref_common.compute_imitation_reward(
        pos[:, 0], rot[:, 0], pos[:, sub], rot[:, sub], vel[:, sub], ang[:, sub],
        ref["rg_pos"][:, sub], ref["rb_rot"][:, sub], ref["body_vel"][:, sub], ref["body_ang_vel"][:, sub], RWD,
    )  # fmt: skip
    arrays["out.reward.subset12"], arrays["out.reward_raw.subset12"] = r.numpy(), raw.numpy()
    # query without offset
    ref2 = lib.get_motion_state(clock.sampled_motion_ids, t, None)
    arrays["out.rg_pos.no_offset"] = ref2["rg_pos"].numpy()
    arrays.update(npify(lib_data.as_dict(), "in.lib"))
    arrays.update(npify(clock.__dict__, "in.clock"))
    arrays["in.t"] = t.numpy()
    save("flags", arrays)


def case_sample_time():
    """MotionLibBase.sample_time_interval (motion_lib.py:526-535) under a fixed torch seed; the
    fixture stores the uniform numbers the same seed produces, so the drop-in takes them as input."""
    lib_data = synth.make_motion_lib(300, 20, 400, fps_choices=(30, 60, 120), seed=77)
    lib = ref_loader.make_reference_lib(lib_data)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 300, (5000,), generator=g)
    torch.manual_seed(4242)
    t = lib.sample_time_interval(ids)
    torch.manual_seed(4242)
    phase = torch.rand(ids.shape)
    save("sample_time", {"in.motion_lengths": lib_data.motion_lengths.numpy(), "in.ids": ids.numpy(),
                         "in.phase": phase.numpy(), "out.motion_time": t.numpy()})  # fmt: skip


def case_amp_obs():
    """build_amp_observations_smpl (common.py:192-267) on seeded inputs: default flags with the
    19-joint dof subset (humanoid_phc.py:186-194), and the flag variants."""
    g = torch.Generator().manual_seed(17)
    n = 96
    root_pos = torch.randn(n, 3, generator=g)
    q = torch.randn(n, 4, generator=g)
    root_rot = q / q.norm(dim=-1, keepdim=True)
    root_vel, root_ang = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    dof_pos = torch.randn(n, 69, generator=g) * torch.exp(torch.rand(n, 1, generator=g) * 12 - 12)  # tiny .. ~1 rad
    dof_pos[::9] = 0
    dof_pos[5::31] *= 6.0  # some angles beyond pi
    dof_vel = torch.randn(n, 69, generator=g)
    key = torch.randn(n, 4, 3, generator=g)
    shape, limb = torch.randn(n, 11, generator=g), torch.randn(n, 10, generator=g)
    removed = (3, 7, 17, 22)  # L_Toe R_Toe L_Hand R_Hand in DOF_NAMES (body_sets.py:42, humanoid_phc.py:186-194)
    subset = torch.tensor([3 * j + k for j in range(23) if j not in removed for k in range(3)])
    arrays = {"in.root_pos": root_pos, "in.root_rot": root_rot, "in.root_vel": root_vel, "in.root_ang_vel": root_ang,
              "in.dof_pos": dof_pos, "in.dof_vel": dof_vel, "in.key_body_pos": key, "in.shape": shape, "in.limb": limb,
              "in.dof_subset": subset}  # fmt: skip
    variants = {  # local_root_obs, root_height_obs, has_dof_subset, has_shape_obs_disc, has_limb_weight_obs, upright
        "default": (True, True, True, False, False, True),
        "all_dofs": (True, True, False, False, False, True),
        "global_root_no_height": (False, False, True, False, False, True),
        "not_upright_with_params": (True, True, True, True, True, False),
    }
    for name, fl in variants.items():
        o = ref_common.build_amp_observations_smpl(root_pos, root_rot, root_vel, root_ang, dof_pos, dof_vel, key, shape,
                                                   limb, subset, *fl)  # fmt: skip
        arrays[f"out.{name}"] = o
        arrays[f"in.flags.{name}"] = torch.tensor(fl)
    save("amp_obs", {k: v.numpy() for k, v in arrays.items()})


def case_running_norm():
    # the policies package __init__ pulls in pufferlib (absent); load the one file directly
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "ref_running_norm", os.path.join(ref_loader.REFERENCE_ROOT, "puffer_phc/policies/running_norm.py")
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    RunningNorm = mod.RunningNorm

    g = torch.Generator().manual_seed(21)
    rn = RunningNorm(934)
    x1 = torch.randn(192, 934, generator=g) * 3 + 1.5
    x2 = torch.randn(160, 934, generator=g) * 0.2 - 4.0
    arrays = {"in.x1": x1.numpy(), "in.x2": x2.numpy()}
    rn.update(x1)
    arrays["out.mean1"], arrays["out.var1"], arrays["out.count1"] = (
        rn.running_mean.numpy().copy(), rn.running_var.numpy().copy(), rn.count.numpy().copy())  # fmt: skip
    rn.update(x2)
    arrays["out.mean2"], arrays["out.var2"], arrays["out.count2"] = (
        rn.running_mean.numpy().copy(), rn.running_var.numpy().copy(), rn.count.numpy().copy())  # fmt: skip
    arrays["out.fwd"] = rn(x2[:32]).numpy()
    save("running_norm", arrays)


def case_episode():
    """PHCPufferEnv.step's per-env episode bookkeeping (clean_pufferl/env.py:109-183), run
    unmodified on a scripted stand-in for HumanoidPHC (isaacgym / pufferlib are absent: both
    imports are stubbed, the bookkeeping code itself is the reference's)."""
    import importlib.util
    import types

    for name, attrs in (
        ("puffer_phc.envs.humanoid_phc", {"HumanoidPHC": object}),
        ("puffer_phc.envs.render_env", {"HumanoidRenderEnv": object}),
        ("pufferlib", {"PufferEnv": type("PufferEnv", (), {})}),
    ):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location(
        "ref_puffer_env", os.path.join(ref_loader.REFERENCE_ROOT, "puffer_phc/clean_pufferl/env.py")
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    N, K, A, LOG = 96, 48, 5, 12
    g = torch.Generator().manual_seed(31)
    rewards = torch.rand(K, N, generator=g) * 1.2
    reward_raw = torch.rand(K, N, 5, generator=g)
    reset = torch.rand(K, N, generator=g) < 0.12
    terminate = reset & (torch.rand(K, N, generator=g) < 0.6)
    reset[7] = False  # a step without any reset takes the other branch
    terminate[7] = False
    actions = (torch.rand(K, N, A, generator=g) * 3 - 1.5).numpy()

    class ScriptedEnv:
        def __init__(self):
            self.obs_buf = torch.zeros(N, 4)
            self.rew_buf = torch.zeros(N)
            self.reset_buf = torch.zeros(N, dtype=torch.bool)
            self.extras = {}
            self.k = 0

        def step(self, actions):
            self.rew_buf[:] = rewards[self.k]
            self.reset_buf[:] = reset[self.k]
            self.extras["terminate"] = terminate[self.k].clone()
            self.extras["reward_raw"] = reward_raw[self.k]
            self.k += 1

        def reset(self, env_ids=None):  # humanoid_phc.py:778-779
            if env_ids is not None:
                self.reset_buf[env_ids] = 0

    env = ScriptedEnv()
    pe = object.__new__(mod.PHCPufferEnv)
    pe.cfg = types.SimpleNamespace(clip_actions=True, use_amp_obs=False, log_interval=LOG, num_envs=N, device="cpu")
    pe.env = env
    bufs = mod.PufferEnvBuffers(env.obs_buf, env.rew_buf, env.reset_buf, N, (A,), "cpu")
    for k in ("observations", "rewards", "terminals", "truncations", "masks", "actions"):
        setattr(pe, k, getattr(bufs, k))
    pe.episode_returns = torch.zeros(N, dtype=torch.float32)
    pe.episode_lengths = torch.zeros(N, dtype=torch.int32)
    pe.episode_count = 0
    pe._infos = {"episode_return": [], "episode_length": [], "truncated_rate": []}
    pe.raw_rewards = torch.zeros(5, dtype=torch.float32)
    pe.tick = 0

    out = {k: [] for k in ("rew", "terminals", "truncations", "masks", "episode_returns", "episode_lengths",
                           "raw_rewards", "episode_count", "actions")}  # fmt: skip
    infos = []
    for k in range(K):
        _, rew, term, trunc, info = pe.step(actions[k])
        out["rew"].append(rew.numpy().copy())
        out["terminals"].append(term.numpy().copy())
        out["truncations"].append(trunc.numpy().copy())
        out["masks"].append(pe.masks.numpy().copy())
        out["episode_returns"].append(pe.episode_returns.numpy().copy())
        out["episode_lengths"].append(pe.episode_lengths.numpy().copy())
        out["raw_rewards"].append(pe.raw_rewards.numpy().copy())
        out["episode_count"].append(pe.episode_count)
        out["actions"].append(pe.actions.numpy().copy())
        if info:
            d = info[0]
            infos.append([k, d["episode_return"], d["episode_length"], d["truncated_rate"], d["rew_body_pos"],
                          d["rew_body_rot"], d["rew_lin_vel"], d["rew_ang_vel"], d["rew_power"]])  # fmt: skip
    arrays = {"in.rewards": rewards.numpy(), "in.reward_raw": reward_raw.numpy(), "in.reset": reset.numpy(),
              "in.terminate": terminate.numpy(), "in.actions": actions, "in.log_interval": np.int64(LOG)}  # fmt: skip
    for k, v in out.items():
        arrays[f"out.{k}"] = np.stack([np.asarray(x) for x in v])
    arrays["out.infos"] = np.asarray(infos, dtype=np.float64)
    save("episode", arrays)


def load_reference_env_module():
    """puffer_phc.envs.humanoid_phc with its simulator imports stubbed (isaacgym, gymtorch, gym are not
    installed and PhysX is out of scope): the class body — step(), reset() and every _compute_* /
    _reset_* / AMP method — is the reference's own code."""
    import types

    for name in ("isaacgym", "isaacgym.gymapi", "isaacgym.gymtorch", "gymtorch", "gym", "gym.spaces"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["isaacgym"].gymapi = sys.modules["isaacgym.gymapi"]
    sys.modules["isaacgym"].gymtorch = sys.modules["isaacgym.gymtorch"]
    sys.modules["gym"].spaces = sys.modules["gym.spaces"]
    for m in ("gymtorch", "isaacgym.gymtorch"):
        sys.modules[m].unwrap_tensor = lambda t: t
    import puffer_phc.envs.humanoid_phc as H

    return H


def make_reference_env(N, S, use_amp=True, state_init=None, seed=41):
    """An instance of the reference's HumanoidPHC built without ``__init__`` (which needs Isaac Gym and asset files),
    with PhysX scripted: ``fetch_results`` writes the next rigid-body / dof state (reference pose at the coming reward
    time + noise) into the same AoS tensors the reference wraps.  Returns (H, env, lib, lib_data, clock, rec, cfg)."""
    import types

    H = load_reference_env_module()
    from puffer_phc.config import EnvConfig

    lib_data, clock, _ = synth.make_case(N, N, ref_query, seed=seed, device="cpu", min_frames=16, max_frames=26,
                                         max_progress=6)  # fmt: skip
    lib = ref_loader.make_reference_lib(lib_data)
    cfg = EnvConfig()
    cfg.num_envs, cfg.device_type, cfg.use_amp_obs, cfg.num_amp_obs_steps = N, "cpu", use_amp, S
    if state_init is not None:
        cfg.state_init = getattr(H.StateInit, state_init)
    assert cfg.device == "cpu" and cfg.robot.freeze_hand and cfg.robot.freeze_toe and cfg.reward.use_power_reward

    env = object.__new__(H.HumanoidPHC)
    env.cfg = cfg
    env.isaac_base = types.SimpleNamespace(dt=DT, control_freq_inv=2)
    env.sim = None
    env.viewer = None
    env.num_dof, env.num_bodies = 69, 24
    env.num_obs = 934
    env._dof_obs_size = 23 * 6
    disc = [np.arange(i * 3, (i + 1) * 3) for i, n in enumerate(H.DOF_NAMES) if n not in H.REMOVE_NAMES]  # :186-194
    env.dof_subset = torch.from_numpy(np.concatenate(disc))
    env._num_amp_obs_per_step = 13 + env._dof_obs_size + env.num_dof + 3 * len(H.KEY_BODIES)  # :470-476
    env._num_amp_obs_per_step -= (6 + 3) * int((env.num_dof - len(env.dof_subset)) / 3)
    env.num_amp_obs = S * env._num_amp_obs_per_step
    env.humanoid_shapes = torch.zeros(N, 17 + 6)
    env.humanoid_limb_and_weights = torch.zeros(N, 10)
    env._humanoid_actor_ids = torch.arange(N, dtype=torch.int32)
    env.all_env_ids = torch.arange(N)
    env._key_body_ids = H.build_body_ids_tensor(H.BODY_NAMES, H.KEY_BODIES, "cpu")
    env._contact_body_ids = H.build_body_ids_tensor(H.BODY_NAMES, H.CONTACT_BODIES, "cpu")
    env._track_bodies_id = H.build_body_ids_tensor(H.BODY_NAMES, H.TRACK_BODIES, "cpu")
    env._reset_bodies_id = H.build_body_ids_tensor(H.BODY_NAMES, H.RESET_BODIES, "cpu")
    env._termination_distances = torch.full((24,), cfg.termination_distance)
    env.flag_im_eval = env.flag_test = False
    env._motion_lib = lib
    # simulator tensors, laid out as gymtorch hands them over (:499-553)
    env._root_states = torch.zeros(N, 13)
    env._humanoid_root_states = env._root_states.view(N, 1, 13)[:, 0]
    env._dof_state = torch.zeros(N * 69, 2)
    env._dof_pos = env._dof_state.view(N, 69, 2)[..., 0]
    env._dof_vel = env._dof_state.view(N, 69, 2)[..., 1]
    env._rigid_body_state = torch.zeros(N * 24, 13)
    rb = env._rigid_body_state.view(N, 24, 13)
    env._rigid_body_pos, env._rigid_body_rot = rb[..., 0:3], rb[..., 3:7]
    env._rigid_body_vel, env._rigid_body_ang_vel = rb[..., 7:10], rb[..., 10:13]
    env.dof_force_tensor = torch.zeros(N, 69)
    env._contact_forces = torch.zeros(N, 24, 3)
    g = torch.Generator().manual_seed(43)
    env._pd_action_offset = torch.randn(69, generator=g) * 0.2
    env._pd_action_scale = torch.rand(69, generator=g) + 0.5
    # env buffers (:556-611), rew_buf as intended there (the torch(...) call at :560 is a typo)
    env.obs_buf = torch.zeros(N, 934)
    env.rew_buf = torch.zeros(N)
    env.reward_raw = torch.zeros(N, 5)
    env.progress_buf = clock.progress_buf.clone()
    env.reset_buf = torch.ones(N, dtype=torch.bool)
    env._terminate_buf = torch.ones(N, dtype=torch.bool)
    env.extras = {}
    env._reset_default_env_ids, env._reset_ref_env_ids = [], []
    env._global_offset = clock.global_offset.clone()
    env._motion_start_times = clock.motion_start_times.clone()
    env._motion_start_times_offset = clock.motion_start_times_offset.clone()
    env._sampled_motion_ids = clock.sampled_motion_ids.clone()
    env.ref_motion_cache = {}
    env._amp_obs_buf = torch.zeros(N, S, env._num_amp_obs_per_step)
    env._curr_amp_obs_buf = env._amp_obs_buf[:, 0]

    class RaisesOnOverlap(torch.Tensor):
        """_update_hist_amp_obs (:1341-1347) assigns buf[:, 0:S-1] to the overlapping view buf[:, 1:] inside a
        try and falls back to .clone() on the RuntimeError the reference's pinned torch 2.3.1 raises there
        (the message is quoted in its comment).  torch 2.11 does not raise and copies front to back, which
        smears slot 0 over the whole history; restore the pinned behaviour so the fallback runs."""

        def __setitem__(self, idx, value):
            if isinstance(value, torch.Tensor) and value.untyped_storage().data_ptr() == self.untyped_storage().data_ptr():
                raise RuntimeError("unsupported operation: some elements of the input tensor and the written-to "
                                   "tensor refer to a single memory location. Please clone() the tensor before "
                                   "performing the operation.")  # fmt: skip
            super().__setitem__(idx, value)

    env._hist_amp_obs_buf = env._amp_obs_buf[:, 1:].as_subclass(RaisesOnOverlap)
    env._amp_obs_demo_buf = torch.zeros_like(env._amp_obs_buf)
    assert torch.equal(env._sampled_motion_ids, torch.arange(N)), "the reference regime: clip i <-> env i"

    rec = {"pd_target": [], "state": [], "dof_state": [], "dof_force": []}

    class ScriptedGym:
        """Every Isaac Gym call is a no-op except the two that matter to the path."""

        def __getattr__(self, name):
            return lambda *a, **k: None

        def set_dof_position_target_tensor(self, sim, t):
            rec["pd_target"].append(t.clone())

        def fetch_results(self, sim, wait):
            # "physics": the pose the reward of THIS step will be compared with, plus noise
            t = (env.progress_buf + 1) * DT + env._motion_start_times + env._motion_start_times_offset
            ref = lib.get_motion_state(env._sampled_motion_ids, t, env._global_offset)
            k = len(rec["state"])
            state = synth.make_sim_state(ref, seed=500 + k)
            env._rigid_body_state.view(N, 24, 13)[:] = state
            env._root_states[:] = state[:, 0]
            gk = torch.Generator().manual_seed(600 + k)
            env._dof_state.view(N, 69, 2)[:] = torch.randn(N, 69, 2, generator=gk)
            env.dof_force_tensor[:] = torch.randn(N, 69, generator=gk) * 30
            rec["state"].append(state.clone())
            rec["dof_state"].append(env._dof_state.view(N, 69, 2).clone())
            rec["dof_force"].append(env.dof_force_tensor.clone())

    env.gym = ScriptedGym()

    return H, env, lib, lib_data, clock, rec, cfg


def case_env_rollout():
    """The unmodified reference ``HumanoidPHC.step(actions)`` and ``HumanoidPHC.reset(env_ids)``
    (envs/humanoid_phc.py:90-172) driven for a few steps as clean_pufferl/env.py:109-140 drives them, on
    an instance built without ``__init__`` (which needs Isaac Gym and asset files).  PhysX is a scripted
    stand-in: ``fetch_results`` writes the next rigid-body / dof state (reference pose at the coming
    reward time + noise) into the same AoS tensors the reference wraps.  Default EnvConfig except
    use_amp_obs=True and num_amp_obs_steps=4 (10 in the config; smaller fixture)."""
    N, K, S = 24, 5, 4
    H, env, lib, lib_data, clock, rec, cfg = make_reference_env(N, S)

    arrays = {}
    arrays.update(npify(lib_data.as_dict(), "in.lib"))
    arrays.update(npify(clock.__dict__, "in.clock"))
    arrays["in.pd_action_offset"], arrays["in.pd_action_scale"] = env._pd_action_offset.numpy(), env._pd_action_scale.numpy()
    arrays["in.dof_subset"] = env.dof_subset.numpy()
    arrays["in.key_body_ids"] = env._key_body_ids.numpy()
    arrays["in.num_amp_obs_steps"] = np.int64(S)
    arrays["in.rew_power_coef"] = np.float32(cfg.rew_power_coef)
    arrays["in.termination_distance"] = np.float32(cfg.termination_distance)
    ga = torch.Generator().manual_seed(44)
    for k in range(K):
        actions = torch.rand(N, 69, generator=ga) * 2 - 1
        arrays[f"in.actions.{k}"] = actions.numpy()
        obs, rew, reset, extras = env.step(actions)
        out = {
            "obs": env.obs_buf, "rew": env.rew_buf, "reward_raw": env.reward_raw, "reset": env.reset_buf,
            "terminate": extras["terminate"], "progress": env.progress_buf, "amp_obs": env.amp_obs,
        }  # fmt: skip
        arrays.update({f"out.step.{k}.{n}": v.detach().numpy().copy() for n, v in out.items()})
        # clean_pufferl/env.py:133-135
        reset_indices = torch.nonzero(env.reset_buf).squeeze(-1)
        arrays[f"in.reset_indices.{k}"] = reset_indices.numpy()
        torch.manual_seed(700 + k)
        arrays[f"in.phase.{k}"] = torch.rand(reset_indices.shape).numpy()  # what sample_time_interval will draw
        torch.manual_seed(700 + k)
        if len(reset_indices) > 0:
            env.reset(reset_indices)
        out = {
            "rigid_body_state": env._rigid_body_state.view(N, 24, 13), "root_states": env._humanoid_root_states,
            "dof_state": env._dof_state.view(N, 69, 2), "obs": env.obs_buf, "progress": env.progress_buf,
            "reset": env.reset_buf, "terminate": env._terminate_buf, "motion_start_times": env._motion_start_times,
            "motion_start_times_offset": env._motion_start_times_offset, "global_offset": env._global_offset,
            "amp_obs": env.amp_obs, "amp_obs_demo": env.fetch_amp_obs_demo(),
        }  # fmt: skip
        arrays.update({f"out.reset.{k}.{n}": v.detach().numpy().copy() for n, v in out.items()})
    for n in ("pd_target", "state", "dof_state", "dof_force"):
        assert len(rec[n]) == K, (n, len(rec[n]))
        for k in range(K):
            arrays[("out" if n == "pd_target" else "in") + f".{n}.{k}"] = rec[n][k].numpy()
    save("env_rollout", arrays)


def case_env_reset_modes():
    """``HumanoidPHC.reset(env_ids)`` with StateInit.Default and StateInit.Hybrid (envs/humanoid_phc.py:678-745):
    the unmodified reference driven like case_env_rollout (AMP off: ``_init_amp_obs`` raises NotImplementedError
    for default-reset envs, :795-797).  Default: root / dof state from the initial buffers, the rigid-body tensors
    and the motion clock untouched, progress 0, observations of the (stale) rigid-body state.  Hybrid:
    ``torch.bernoulli(hybrid_init_prob)`` per env picks reference-state init or default; the mask and the uniform
    numbers ``sample_time_interval`` draws afterwards are recorded."""
    N, K, S = 24, 3, 4
    arrays = {}
    for mode in ("Default", "Hybrid"):
        H, env, lib, lib_data, clock, rec, cfg = make_reference_env(N, S, use_amp=False, state_init=mode, seed=47)
        cfg.hybrid_init_prob = 0.5
        g = torch.Generator().manual_seed(48)
        env._initial_humanoid_root_states = torch.randn(N, 13, generator=g)
        env._initial_dof_pos = torch.randn(N, 69, generator=g) * 0.3
        env._initial_dof_vel = torch.randn(N, 69, generator=g)
        if mode == "Default":  # shared inputs, once
            arrays.update(npify(lib_data.as_dict(), "in.lib"))
            arrays.update(npify(clock.__dict__, "in.clock"))
            arrays["in.pd_action_offset"], arrays["in.pd_action_scale"] = env._pd_action_offset.numpy(), env._pd_action_scale.numpy()
            arrays["in.rew_power_coef"] = np.float32(cfg.rew_power_coef)
            arrays["in.termination_distance"] = np.float32(cfg.termination_distance)
            arrays["in.hybrid_init_prob"] = np.float32(cfg.hybrid_init_prob)
            arrays["in.initial_root_states"] = env._initial_humanoid_root_states.numpy()
            arrays["in.initial_dof_pos"] = env._initial_dof_pos.numpy()
            arrays["in.initial_dof_vel"] = env._initial_dof_vel.numpy()
        ga = torch.Generator().manual_seed(49)
        for k in range(K):
            actions = torch.rand(N, 69, generator=ga) * 2 - 1
            arrays[f"in.{mode}.actions.{k}"] = actions.numpy()
            obs, rew, reset, extras = env.step(actions)
            out = {"obs": env.obs_buf, "rew": env.rew_buf, "reward_raw": env.reward_raw, "reset": env.reset_buf,
                   "terminate": extras["terminate"], "progress": env.progress_buf}  # fmt: skip
            arrays.update({f"out.{mode}.step.{k}.{n}": v.detach().numpy().copy() for n, v in out.items()})
            reset_indices = torch.nonzero(env.reset_buf).squeeze(-1)
            assert 0 < len(reset_indices) < N
            arrays[f"in.{mode}.reset_indices.{k}"] = reset_indices.numpy()
            torch.manual_seed(800 + k)
            if mode == "Hybrid":  # what _reset_hybrid_state_init (:733-745) and then sample_time_interval will draw
                probs = H.to_torch(np.array([cfg.hybrid_init_prob] * len(reset_indices)), device="cpu")
                ref_mask = torch.bernoulli(probs) == 1.0
                arrays[f"in.{mode}.ref_mask.{k}"] = ref_mask.numpy()
                arrays[f"in.{mode}.phase.{k}"] = torch.rand(int(ref_mask.sum())).numpy()
                assert 0 < int(ref_mask.sum()) < len(reset_indices)
            torch.manual_seed(800 + k)
            env.reset(reset_indices)
            out = {
                "rigid_body_state": env._rigid_body_state.view(N, 24, 13), "root_states": env._humanoid_root_states,
                "dof_state": env._dof_state.view(N, 69, 2), "obs": env.obs_buf, "progress": env.progress_buf,
                "reset": env.reset_buf, "terminate": env._terminate_buf, "motion_start_times": env._motion_start_times,
                "motion_start_times_offset": env._motion_start_times_offset, "global_offset": env._global_offset,
            }  # fmt: skip
            arrays.update({f"out.{mode}.reset.{k}.{n}": v.detach().numpy().copy() for n, v in out.items()})
        for n in ("state", "dof_state", "dof_force"):
            for k in range(K):
                arrays[f"in.{mode}.{n}.{k}"] = rec[n][k].numpy()
        for k in range(K):
            arrays[f"out.{mode}.pd_target.{k}"] = rec["pd_target"][k].numpy()
    save("env_reset_modes", arrays)


def case_pd_offset_scale():
    """``HumanoidPHC._build_pd_action_offset_scale`` (envs/humanoid_phc.py:385-457), the reference's own method run on
    an instance built without ``__init__``: what ``_action_to_pd_targets`` scales the actions with, from the asset's
    joint limits.  All four (bias_offset, has_smpl_pd_offset x has_upright_start) variants."""
    import types

    H = load_reference_env_module()
    g = torch.Generator().manual_seed(61)
    lo = -(torch.rand(69, generator=g) * 2.5 + 0.1)
    hi = torch.rand(69, generator=g) * 2.5 + 0.1
    arrays = {"in.dof_limits_lower": lo.numpy(), "in.dof_limits_upper": hi.numpy()}
    for name, bias, smpl_off, upright in (("default", False, False, True), ("bias_offset", True, False, True),
                                          ("smpl_pd_offset_upright", False, True, True),
                                          ("smpl_pd_offset_not_upright", False, True, False)):  # fmt: skip
        env = object.__new__(H.HumanoidPHC)
        env.cfg = types.SimpleNamespace(device="cpu", robot=types.SimpleNamespace(
            bias_offset=bias, has_smpl_pd_offset=smpl_off, has_upright_start=upright))
        env.dof_limits_lower, env.dof_limits_upper = lo.clone(), hi.clone()
        env._dof_offsets = np.linspace(0, 69, 24).astype(int)  # :220
        env._build_pd_action_offset_scale()
        arrays[f"out.{name}.offset"] = env._pd_action_offset.numpy()
        arrays[f"out.{name}.scale"] = env._pd_action_scale.numpy()
        arrays[f"in.{name}"] = np.asarray([bias, smpl_off, upright])
    save("pd_offset_scale", arrays)


def synth_build_clips(tree_parents, seed=5, shapes=((20, 30), (45, 30), (3, 30), (33, 60), (12, 30), (2, 30))):
    """Synthetic stand-ins for the AMASS pkl entries (scripts/phc_convert_amass_data.py:186-194):
    per clip ``root_trans_offset`` (torch f64), ``pose_aa`` (numpy f64 [B,72]), ``pose_quat_global``
    (numpy f64 [B,24,4], a smooth rotation per joint composed down the SMPL tree), ``fps``."""
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(seed)
    clips = {}
    for i, (B, fps) in enumerate(shapes):
        base = R.random(24, random_state=100 + i)
        w = rng.normal(size=(24, 3)) * 0.8
        quat = np.zeros((B, 24, 4))
        for f in range(B):
            g = [None] * 24
            for j in range(24):
                loc = R.from_rotvec(w[j] * f / fps) * base[j]
                g[j] = loc if tree_parents[j] < 0 else g[tree_parents[j]] * loc
                quat[f, j] = g[j].as_quat()
        if i == 4:  # quaternions that went through fp32 storage: unit only to 1e-8
            quat = quat.astype(np.float32).astype(np.float64)
        trans = np.cumsum(rng.normal(size=(B, 3)) * 0.02, 0) + np.array([0.3 * i, -0.2 * i, 0.9])
        pose_aa = rng.normal(size=(B, 72)) * 0.7
        pose_aa[:: 5, :3] *= 1e-4  # small-angle branch of from_rotvec / as_rotvec
        clip = {"root_trans_offset": torch.from_numpy(trans), "pose_aa": pose_aa, "pose_quat_global": quat,
                "beta": np.zeros(16), "gender": "neutral"}  # fmt: skip
        if fps != 30:
            clip["fps"] = fps  # the loader defaults to 30 (motion_lib.py:806)
        clips[f"clip{i}"] = clip
    return clips


def case_motion_build():
    """The reference's own ``MotionLibSMPL.load_motions`` (PHC/motion_lib.py:257-428, worker
    ``load_motion_with_skeleton`` :748-824: forward kinematics and velocity estimation through
    ``SkeletonState`` / ``SkeletonMotion`` of PHC/poselib_skeleton.py, ``compute_motion_dof_vels_jit`` :120)
    on synthetic clips and the SMPL skeleton tree of the reference's MJCF asset.  ``__init__`` needs an AMASS
    pkl + SMPL model files, so the object is made with ``object.__new__`` and the attributes ``load_data`` /
    ``setup_constants`` would set; ``mesh_parsers`` is None as when the SMPL models are not found (:692-694).
    Two variants: deterministic (no heading randomisation) and the training default (random heading per
    clip, :789-799) under a fixed numpy seed; the fixture stores the uniform numbers that seed produces."""
    import types

    from puffer_phc.poselib_skeleton import SkeletonTree

    tree = SkeletonTree.from_mjcf(os.path.join(ref_loader.REFERENCE_ROOT, "puffer_phc/assets/smpl_humanoid.xml"))
    parents = tree.parent_indices.numpy()
    g = torch.Generator().manual_seed(9)
    # a second body shape: the env holds one tree per humanoid (humanoid_phc.py:208-209, :360)
    tree2 = SkeletonTree(tree.node_names, tree.parent_indices,
                         tree.local_translation * (1.0 + 0.1 * torch.rand(24, 3, generator=g)))  # fmt: skip
    clips = synth_build_clips(parents)
    M = len(clips)
    trees = [tree if i % 2 == 0 else tree2 for i in range(M)]
    gender_betas = [torch.cat([torch.zeros(1), torch.randn(16, generator=g)]) for _ in range(M)]
    limb_weights = [np.random.default_rng(i).normal(size=10) for i in range(M)]

    arrays = {
        "in.parent_indices": parents.astype(np.int32),
        "in.local_translation": np.stack([t.local_translation.numpy() for t in trees]),
        "in.pose_quat_global": np.concatenate([c["pose_quat_global"] for c in clips.values()]),
        "in.root_trans_offset": np.concatenate([c["root_trans_offset"].numpy() for c in clips.values()]),
        "in.pose_aa": np.concatenate([c["pose_aa"] for c in clips.values()]).copy(),
        "in.num_frames": np.asarray([c["pose_aa"].shape[0] for c in clips.values()], dtype=np.int64),
        "in.fps": np.asarray([c.get("fps", 30) for c in clips.values()], dtype=np.int64),
        "in.gender_betas": torch.stack(gender_betas).numpy(),
        "in.limb_weights": np.stack(limb_weights),
    }

    def run(deterministic, seed, max_length=-1, pyseed=0):
        import copy
        import random

        data = copy.deepcopy(clips)  # the loader rotates pose_aa[:, :3] of its input in place (:794)
        lib = object.__new__(ref_ml.MotionLibSMPL)
        lib.m_cfg = types.SimpleNamespace(max_length=max_length, fix_height=ref_ml.FixHeightMode.no_fix,
                                          is_deterministic=deterministic, im_eval=False, num_thread=1)  # fmt: skip
        lib._device, lib.mesh_parsers, lib.num_thread = "cpu", None, 1
        lib._motion_data_list = np.array(list(data.values()))
        lib._motion_data_keys = np.array(list(data.keys()))
        lib._num_unique_motions = M
        lib._sampling_prob = torch.ones(M) / M
        np.random.seed(seed)
        random.seed(pyseed)  # the crop start of clips longer than max_length (:776)
        lib.load_motions(trees, gender_betas, limb_weights, random_sample=False)
        out = {k: getattr(lib, k) for k in (
            "gts", "grs", "lrs", "gvs", "gavs", "dvs", "grvs", "gravs", "_motion_aa", "_motion_lengths",
            "_motion_num_frames", "_motion_dt", "_motion_fps", "length_starts", "_motion_bodies",
            "_motion_limb_weights")}  # fmt: skip
        return {k.lstrip("_"): v.numpy() for k, v in out.items()}

    for k, v in run(True, 0).items():
        arrays[f"out.deterministic.{k}"] = v
    for k, v in run(False, 31).items():
        arrays[f"out.random_heading.{k}"] = v
    np.random.seed(31)
    arrays["in.heading_u"] = np.asarray([np.random.random() for _ in range(M)])
    # third variant: max_length crop at a random start (python's `random`) + random heading, the training default
    # with cfg.max_length set; the draws the loader makes, in its order, are recorded for the oracle
    import random

    CROP = 16
    for k, v in run(False, 31, max_length=CROP, pyseed=7).items():
        arrays[f"out.cropped.{k}"] = v
    random.seed(7)
    lens = arrays["in.num_frames"].tolist()
    arrays["in.crop_start"] = np.asarray([0 if n < CROP else random.randint(0, n - CROP) for n in lens], dtype=np.int64)
    arrays["in.crop_max_length"] = np.asarray(CROP, dtype=np.int64)
    save("motion_build", arrays)


if __name__ == "__main__":
    torch.set_num_threads(1)
    torch.manual_seed(0)
    if len(sys.argv) > 1:  # regenerate only the named fixtures, e.g. `make_golden.py episode`
        for name in sys.argv[1:]:
            globals()[f"case_{name}"]()
        sys.exit(0)
    # the reference regime: ids == arange, 30 fps clips, frame-aligned starts
    case_step("step_aligned", N=32, M=32, seed=11, min_frames=8, max_frames=16, max_progress=10)
    # config-4 style: random ids over a shared library, mixed fps, unaligned times, i.i.d. rotations
    case_step("step_random", N=64, M=7, seed=12, min_frames=10, max_frames=40, fps_choices=(30, 60, 120),
              ids="random", aligned=False, rot_regime="random", max_progress=6)  # fmt: skip
    # config-5 style: 10 future frames
    case_step("step_T10", N=24, M=6, seed=13, time_steps=10, min_frames=20, max_frames=40, max_progress=20)
    # eval-mode reset: mean over the 20 EVAL_BODIES (humanoid_phc.py:1431-1437); the threshold is
    # 0.15 m here instead of 0.5 m so that the synthetic noise produces both outcomes
    case_step("step_eval_reset", N=48, M=6, seed=14, min_frames=10, max_frames=30, reset_body_ids=EVAL_BODY_IDS,
              use_mean=True, term=0.15, max_progress=6)  # fmt: skip
    case_step("step_no_early_term", N=16, M=4, seed=15, min_frames=6, max_frames=12, early=False, max_progress=14)
    case_frame_blend()
    case_slerp()
    case_flags()
    case_running_norm()
    case_sample_time()
    case_reset_mean()
    case_amp_obs()
    case_episode()
    case_motion_build()
    case_env_rollout()
    case_env_reset_modes()
    case_pd_offset_scale()
