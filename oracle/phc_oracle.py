"""CPU oracle for the PHC step path.  TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch-CPU fp32 array ops, of the reference algorithm for the
path BASELINE.json names: motion-state query, smpl_max self observation, imitation
observation v6 (and the v7 column subset), imitation reward, reset/termination, the
step orchestration around them and the RunningNorm update.  Each function cites the
reference lines it follows (paths relative to
/root/reference/packages/puffer-phc/puffer_phc/).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module, and only as the checker / the timed CPU baseline.  The product
(``humanoid_b200``) never imports it and has no CPU fallback.

Parity pin: ``tests/golden/make_golden.py`` runs the *reference's own functions*
(imported from /root/reference in the build container) on seeded inputs and commits
inputs + outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this module against them: bit-exact for frame indices and reset flags, and (because it
issues the same ATen CPU ops in the same order) bit-exact for the floats as well when
run on the same torch build that generated the fixtures — a 1e-6 tolerance is used so
that a different host ISA (AVX2 vs AVX-512 reductions) cannot break the pin.

Rounding facts the port relies on, measured against the reference on torch 2.11 CPU
(see DESIGN.md "oracle notes"): ``sum(dim=-1)`` over 4 is ((a+b)+c)+d; ``bmm`` dot of
3 is sequential without FMA; ``norm(dim=-1)`` over 3 is sqrt(fma(z,z,fma(y,y,x*x)));
``cross`` is fma(a1,b2,-(a2*b1)); ``mean`` is sum/count.  The port calls the same
ATen ops for those, so it inherits them instead of restating them.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------
# quaternion primitives (xyzw) — torch_utils.py
# ----------------------------------------------------------------------------------


def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    """8-multiply product, torch_utils.py:55-75 (same grouping of terms)."""
    x1, y1, z1, w1 = a.unbind(-1)
    x2, y2, z2, w2 = b.unbind(-1)
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack((x, y, z, w), dim=-1)


def quat_conj(q: Tensor) -> Tensor:
    """torch_utils.py:79-82."""
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def rotate(q: Tensor, v: Tensor) -> Tensor:
    """my_quat_rotate, torch_utils.py:274-281: (v(2w^2-1) + 2w(q x v)) + 2q(q.v).

    ``q`` and ``v`` broadcast against each other; the reference flattens to [n,4]/[n,3]
    and uses torch.cross (FMA inside) and bmm (sequential, no FMA) — the dot below is
    the same three separately-rounded products added left to right."""
    qw = q[..., 3:4]
    qv = q[..., :3]
    shape = torch.broadcast_shapes(qv.shape, v.shape)
    qv_e = qv.expand(shape)
    v_e = v.expand(shape)
    a = v_e * (2.0 * qw**2 - 1.0)
    b = torch.linalg.cross(qv_e, v_e, dim=-1) * qw * 2.0
    dot = (qv_e[..., 0:1] * v_e[..., 0:1] + qv_e[..., 1:2] * v_e[..., 1:2]) + qv_e[..., 2:3] * v_e[..., 2:3]
    c = qv_e * dot * 2.0
    return a + b + c


def tan_norm(q: Tensor) -> Tensor:
    """quat_to_tan_norm, torch_utils.py:285-297: rotated x axis then rotated z axis."""
    ex = torch.zeros_like(q[..., :3])
    ex[..., 0] = 1
    ez = torch.zeros_like(q[..., :3])
    ez[..., 2] = 1
    return torch.cat((rotate(q, ex), rotate(q, ez)), dim=-1)


def heading_angle(q: Tensor) -> Tensor:
    """calc_heading, torch_utils.py:369-380."""
    ex = torch.zeros_like(q[..., :3])
    ex[..., 0] = 1
    d = rotate(q, ex)
    return torch.atan2(d[..., 1], d[..., 0])


def quat_from_z_angle(angle: Tensor) -> Tensor:
    """quat_from_angle_axis with axis (0,0,1), torch_utils.py:354-358 (+ normalize :45,
    quat_unit :174-179)."""
    axis = torch.zeros(angle.shape + (3,), dtype=angle.dtype, device=angle.device)
    axis[..., 2] = 1
    half = (angle / 2).unsqueeze(-1)
    unit_axis = axis / axis.norm(p=2, dim=-1).clamp(min=1e-9, max=None).unsqueeze(-1)
    q = torch.cat((unit_axis * half.sin(), half.cos()), dim=-1)
    return q / q.norm(p=2, dim=-1).unsqueeze(-1).clamp(min=1e-9)


def heading_quat(q: Tensor) -> Tensor:
    """calc_heading_quat, torch_utils.py:384-394."""
    return quat_from_z_angle(heading_angle(q))


def heading_quat_inv(q: Tensor) -> Tensor:
    """calc_heading_quat_inv, torch_utils.py:398-408."""
    return quat_from_z_angle(-heading_angle(q))


def angle_axis(q: Tensor) -> Tuple[Tensor, Tensor]:
    """quat_to_angle_axis, torch_utils.py:86-106."""
    w = q[..., 3]
    s = torch.sqrt(1 - w * w)
    ang = 2 * torch.acos(w)
    ang = torch.atan2(torch.sin(ang), torch.cos(ang))  # normalize_angle :50-51
    axis = q[..., 0:3] / s.unsqueeze(-1)
    ok = torch.abs(s) > 1e-5
    fallback = torch.zeros_like(axis)
    fallback[..., 2] = 1
    ang = torch.where(ok, ang, torch.zeros_like(ang))
    axis = torch.where(ok.unsqueeze(-1), axis, fallback)
    return ang, axis


def exp_map(q: Tensor) -> Tensor:
    """quat_to_exp_map, torch_utils.py:135-150."""
    ang, axis = angle_axis(q)
    return ang.unsqueeze(-1) * axis


def slerp(q0: Tensor, q1: Tensor, t: Tensor) -> Tensor:
    """torch_utils.py:110-131.  Not renormalised; the two guards are applied in the
    reference's order so the |cos|>=1 select wins."""
    c = torch.sum(q0 * q1, dim=-1)
    flip = c < 0
    q1 = torch.where(flip.unsqueeze(-1), -q1, q1)
    c = torch.abs(c).unsqueeze(-1)
    th = torch.acos(c)
    s = torch.sqrt(1.0 - c * c)
    ra = torch.sin((1 - t) * th) / s
    rb = torch.sin(t * th) / s
    out = ra * q0 + rb * q1
    out = torch.where(torch.abs(s) < 0.001, 0.5 * q0 + 0.5 * q1, out)
    out = torch.where(torch.abs(c) >= 1, q0, out)
    return out


_BASE_ROT_CONJ = (-0.5, -0.5, -0.5, 0.5)


def remove_base_rot(q: Tensor) -> Tensor:
    """envs/common.py:15-19."""
    base = torch.tensor(_BASE_ROT_CONJ, dtype=q.dtype, device=q.device).expand_as(q)
    return quat_mul(q, base)


# ----------------------------------------------------------------------------------
# motion library query — motion_lib.py
# ----------------------------------------------------------------------------------


def frame_blend(time: Tensor, length: Tensor, num_frames: Tensor, dt: Tensor):
    """_calc_frame_blend, motion_lib.py:655-665.  ``phase`` is formed from the
    un-clamped time; negative times are zeroed afterwards."""
    time = time.clone()
    phase = torch.clip(time / length, 0.0, 1.0)
    time[time < 0] = 0
    idx0 = (phase * (num_frames - 1)).long()
    idx1 = torch.min(idx0 + 1, num_frames - 1)
    blend = torch.clip((time - idx0 * dt) / dt, 0.0, 1.0)
    return idx0, idx1, blend


class OracleMotionLib:
    """Holds the A0 tensor set under the reference's attribute names
    (motion_lib.py:396-420) and answers ``get_motion_state`` (motion_lib.py:549-626)."""

    def __init__(self, data):
        d = data.as_dict() if hasattr(data, "as_dict") else dict(data)
        self.gts, self.grs, self.lrs = d["gts"], d["grs"], d["lrs"]
        self.gvs, self.gavs, self.dvs = d["gvs"], d["gavs"], d["dvs"]
        self._motion_aa = d["motion_aa"]
        self._motion_lengths = d["motion_lengths"]
        self._motion_num_frames = d["motion_num_frames"]
        self._motion_dt = d["motion_dt"]
        self.length_starts = d["length_starts"]
        self._motion_bodies = d["motion_bodies"]
        self._motion_limb_weights = d["motion_limb_weights"]

    def _calc_frame_blend(self, time, length, num_frames, dt):
        return frame_blend(time, length, num_frames, dt)

    def get_motion_state(self, motion_ids: Tensor, motion_times: Tensor, offset: Optional[Tensor] = None):
        idx0, idx1, blend = frame_blend(
            motion_times,
            self._motion_lengths[motion_ids],
            self._motion_num_frames[motion_ids],
            self._motion_dt[motion_ids],
        )
        base = self.length_starts[motion_ids]
        f0, f1 = idx0 + base, idx1 + base
        b = blend.view(-1, 1, 1)

        def lerp(table):
            return (1.0 - b) * table[f0] + b * table[f1]

        rg_pos = lerp(self.gts)
        if offset is not None:
            rg_pos = rg_pos + offset[:, None, :]
        body_vel = lerp(self.gvs)
        body_ang_vel = lerp(self.gavs)
        dof_vel = lerp(self.dvs)
        local_rot = slerp(self.lrs[f0], self.lrs[f1], b)
        rb_rot = slerp(self.grs[f0], self.grs[f1], b)
        dof_pos = exp_map(local_rot[:, 1:]).reshape(local_rot.shape[0], -1)  # :670-673
        return {
            "root_pos": rg_pos[:, 0].clone(),
            "root_rot": rb_rot[:, 0].clone(),
            "dof_pos": dof_pos,
            "root_vel": body_vel[:, 0].clone(),
            "root_ang_vel": body_ang_vel[:, 0].clone(),
            "dof_vel": dof_vel.reshape(dof_vel.shape[0], -1),
            "motion_aa": self._motion_aa[f0],
            "rg_pos": rg_pos,
            "rb_rot": rb_rot,
            "body_vel": body_vel,
            "body_ang_vel": body_ang_vel,
            "motion_bodies": self._motion_bodies[motion_ids],
            "motion_limb_weights": self._motion_limb_weights[motion_ids],
            # extras (not in the reference dict) so tests can pin the integer path
            "frame_idx0": idx0,
            "frame_idx1": idx1,
            "blend": blend,
        }


# ----------------------------------------------------------------------------------
# observations — envs/common.py
# ----------------------------------------------------------------------------------


def self_obs_smpl_max(
    body_pos: Tensor,
    body_rot: Tensor,
    body_vel: Tensor,
    body_ang_vel: Tensor,
    smpl_params: Optional[Tensor],
    limb_weight_params: Optional[Tensor],
    local_root_obs: bool,
    root_height_obs: bool,
    upright: bool,
    has_smpl_params: bool,
    has_limb_weight_params: bool,
) -> Tensor:
    """compute_humanoid_observations_smpl_max, envs/common.py:23-103."""
    n = body_pos.shape[0]
    root_pos = body_pos[:, 0]
    root_rot = body_rot[:, 0]
    if not upright:
        root_rot = remove_base_rot(root_rot)
    hinv = heading_quat_inv(root_rot)[:, None, :]

    rel = rotate(hinv, body_pos - root_pos[:, None, :]).reshape(n, -1)[:, 3:]
    rot6 = tan_norm(quat_mul(hinv.expand_as(body_rot), body_rot)).reshape(n, -1).clone()
    if not local_root_obs:
        rot6[:, 0:6] = tan_norm(root_rot)
    vel = rotate(hinv, body_vel).reshape(n, -1)
    ang = rotate(hinv, body_ang_vel).reshape(n, -1)

    parts = []
    if root_height_obs:
        parts.append(root_pos[:, 2:3])
    parts += [rel, rot6, vel, ang]
    if has_smpl_params:
        parts.append(smpl_params)
    if has_limb_weight_params:
        parts.append(limb_weight_params)
    return torch.cat(parts, dim=-1)


def imitation_obs_v6(
    root_pos: Tensor,
    root_rot: Tensor,
    body_pos: Tensor,
    body_rot: Tensor,
    body_vel: Tensor,
    body_ang_vel: Tensor,
    ref_body_pos: Tensor,
    ref_body_rot: Tensor,
    ref_body_vel: Tensor,
    ref_body_ang_vel: Tensor,
    time_steps: int,
    upright: bool,
) -> Tensor:
    """compute_imitation_observations_v6, envs/common.py:106-176.  Reference-state
    inputs are [B*T, J, .] env-major then t; output [B, T*J*24], t-major blocks."""
    B, J, _ = body_pos.shape
    T = time_steps
    if not upright:
        root_rot = remove_base_rot(root_rot)
    hinv = heading_quat_inv(root_rot).view(B, 1, 1, 4)
    h = heading_quat(root_rot).view(B, 1, 1, 4)

    rp = ref_body_pos.reshape(B, T, J, 3)
    rr = ref_body_rot.reshape(B, T, J, 4)
    rv = ref_body_vel.reshape(B, T, J, 3)
    ra = ref_body_ang_vel.reshape(B, T, J, 3)
    hinv_q = hinv.expand(B, T, J, 4)
    h_q = h.expand(B, T, J, 4)

    d_pos = rotate(hinv, rp - body_pos.reshape(B, 1, J, 3))
    d_rot_g = quat_mul(rr, quat_conj(body_rot.reshape(B, 1, J, 4)).expand(B, T, J, 4))
    d_rot = tan_norm(quat_mul(quat_mul(hinv_q, d_rot_g), h_q))
    d_vel = rotate(hinv, rv - body_vel.reshape(B, 1, J, 3))
    d_ang = rotate(hinv, ra - body_ang_vel.reshape(B, 1, J, 3))
    l_pos = rotate(hinv, rp - root_pos.reshape(B, 1, 1, 3))
    l_rot = tan_norm(quat_mul(hinv_q, rr))

    blocks = [x.reshape(B, T, -1) for x in (d_pos, d_rot, d_vel, d_ang, l_pos, l_rot)]
    return torch.cat(blocks, dim=-1).reshape(B, -1)


def v7_columns(num_bodies: int, time_steps: int) -> Tensor:
    """Column indices of the v7 subset inside a v6 row: per future
    step [d_pos | d_vel | l_pos].  The reference has no v7; this is the slicing oracle."""
    J = num_bodies
    per_t = 24 * J
    cols = []
    for t in range(time_steps):
        base = t * per_t
        cols += list(range(base, base + 3 * J))
        cols += list(range(base + 9 * J, base + 12 * J))
        cols += list(range(base + 15 * J, base + 18 * J))
    return torch.tensor(cols, dtype=torch.int64)


def imitation_obs_v7(*args) -> Tensor:
    """v7 = v6[:, v7_columns] — parity against /root/reference is unpinned (no v7 there)."""
    v6 = imitation_obs_v6(*args)
    J, T = args[2].shape[1], args[10]
    return v6[:, v7_columns(J, T)]


# ----------------------------------------------------------------------------------
# AMP observations — envs/common.py:179-267, torch_utils.py:334-366
# ----------------------------------------------------------------------------------


def exp_map_to_quat(em: Tensor) -> Tensor:
    """exp_map_to_angle_axis + quat_from_angle_axis, torch_utils.py:334-366."""
    ang = torch.norm(em, dim=-1)
    axis = em / ang.unsqueeze(-1)
    ang = torch.atan2(torch.sin(ang), torch.cos(ang))
    ok = torch.abs(ang) > 1e-5
    fallback = torch.zeros_like(em)
    fallback[..., 2] = 1
    ang = torch.where(ok, ang, torch.zeros_like(ang))
    axis = torch.where(ok.unsqueeze(-1), axis, fallback)
    half = (ang / 2).unsqueeze(-1)
    unit_axis = axis / axis.norm(p=2, dim=-1).clamp(min=1e-9, max=None).unsqueeze(-1)
    q = torch.cat((unit_axis * half.sin(), half.cos()), dim=-1)
    return q / q.norm(p=2, dim=-1).unsqueeze(-1).clamp(min=1e-9)


def amp_obs_smpl(
    root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, shape_params, limb_weight_params,
    dof_subset, local_root_obs: bool, root_height_obs: bool, has_dof_subset: bool, has_shape_obs_disc: bool,
    has_limb_weight_obs: bool, upright: bool,
) -> Tensor:  # fmt: skip
    """build_amp_observations_smpl, envs/common.py:192-267 (dof_to_obs_smpl :179-189)."""
    B = root_pos.shape[0]
    root_h = root_pos[:, 2:3]
    if not upright:
        root_rot = remove_base_rot(root_rot)
    hinv = heading_quat_inv(root_rot)
    root_rot_obs = tan_norm(quat_mul(hinv, root_rot) if local_root_obs else root_rot)
    local_root_vel = rotate(hinv, root_vel)
    local_root_ang_vel = rotate(hinv, root_ang_vel)
    local_key = rotate(hinv[:, None, :], key_body_pos - root_pos[:, None, :]).reshape(B, -1)
    if has_dof_subset:
        dof_vel = dof_vel[:, dof_subset]
        dof_pos = dof_pos[:, dof_subset]
    dof_obs = tan_norm(exp_map_to_quat(dof_pos.reshape(-1, 3))).reshape(B, -1)
    parts = [root_h] if root_height_obs else []
    parts += [root_rot_obs, local_root_vel, local_root_ang_vel, dof_obs, dof_vel, local_key]
    if has_shape_obs_disc:
        parts.append(shape_params)
    if has_limb_weight_obs:
        parts.append(limb_weight_params)
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------------
# reward and reset — envs/common.py
# ----------------------------------------------------------------------------------


def imitation_reward(
    root_pos,
    root_rot,
    body_pos,
    body_rot,
    body_vel,
    body_ang_vel,
    ref_body_pos,
    ref_body_rot,
    ref_body_vel,
    ref_body_ang_vel,
    rwd_specs: Dict[str, float],
):
    """compute_imitation_reward, envs/common.py:270-322."""
    k = [float(rwd_specs[n]) for n in ("k_pos", "k_rot", "k_vel", "k_ang_vel")]
    w = [float(rwd_specs[n]) for n in ("w_pos", "w_rot", "w_vel", "w_ang_vel")]

    def msq(d):
        return (d**2).mean(dim=-1).mean(dim=-1)

    d_pos = msq(ref_body_pos - body_pos)
    ang, _ = angle_axis(quat_mul(ref_body_rot, quat_conj(body_rot)))
    d_rot = (ang**2).mean(dim=-1)
    d_vel = msq(ref_body_vel - body_vel)
    d_ang = msq(ref_body_ang_vel - body_ang_vel)
    r = [torch.exp(-k[i] * d) for i, d in enumerate((d_pos, d_rot, d_vel, d_ang))]
    reward = w[0] * r[0] + w[1] * r[1] + w[2] * r[2] + w[3] * r[3]
    return reward, torch.stack(r, dim=-1)


def im_reset(
    reset_buf,
    progress_buf,
    contact_buf,
    contact_body_ids,
    rigid_body_pos,
    ref_body_pos,
    pass_time,
    enable_early_termination: bool,
    termination_distance,
    use_mean: bool,
):
    """compute_humanoid_im_reset, envs/common.py:325-364 (contact args never read)."""
    terminated = torch.zeros_like(reset_buf)
    if enable_early_termination:
        dist = torch.norm(rigid_body_pos - ref_body_pos, dim=-1)
        if use_mean:
            fallen = torch.any(dist.mean(dim=-1, keepdim=True) > termination_distance[0], dim=-1)
        else:
            fallen = torch.any(dist > termination_distance, dim=-1)
        fallen = fallen * (progress_buf > 1)
        terminated = torch.where(fallen, torch.ones_like(reset_buf), terminated)
    reset = torch.where(pass_time, torch.ones_like(reset_buf), terminated)
    return reset, terminated


# ----------------------------------------------------------------------------------
# step orchestration — envs/humanoid_phc.py:138-149
# ----------------------------------------------------------------------------------

DEFAULT_RWD_SPECS = dict(  # config.py:38-50 (asdict(RewardConfig); extra keys are ignored)
    k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1,
    w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1,
    imitation_reward_dim=4, full_body_reward=True, use_power_reward=True,
)  # fmt: skip


def step(
    lib: OracleMotionLib,
    state: Tensor,  # [N, bodies_per_env, 13] AoS sim state (after the physics step)
    progress_buf: Tensor,  # [N] int16 — incremented in place, as step() does (:138)
    motion_start_times: Tensor,
    motion_start_times_offset: Tensor,
    global_offset: Tensor,
    sampled_motion_ids: Tensor,
    termination_distances: Tensor,  # [24]
    dt: float,
    reset_buf: Optional[Tensor] = None,
    reset_body_ids: Optional[Tensor] = None,
    use_mean: bool = False,
    enable_early_termination: bool = True,
    rwd_specs: Optional[Dict[str, float]] = None,
    time_steps: int = 1,
    num_bodies: int = 24,
    dof_force: Optional[Tensor] = None,
    dof_vel: Optional[Tensor] = None,
    rew_power_coef: float = 0.0005,
):
    """post-physics half of HumanoidPHC.step (humanoid_phc.py:138-149):
    progress += 1; _compute_reward (:1230-1271); _compute_reset (:1313-1335);
    _compute_observations (:937-961 -> :963-998, :1050-1123).  Returns
    obs [N, 358+576*T], reward [N], reward_raw [N,4], reset [N] bool, terminated [N] bool.
    For T>1 the future reference frames are queried at t+dt .. t+T*dt (config 5)."""
    rwd_specs = rwd_specs or DEFAULT_RWD_SPECS
    N = state.shape[0]
    J = num_bodies
    pos = state[:, :J, 0:3]
    rot = state[:, :J, 3:7]
    vel = state[:, :J, 7:10]
    ang = state[:, :J, 10:13]
    if reset_buf is None:
        reset_buf = torch.ones(N, dtype=torch.bool, device=state.device)
    if reset_body_ids is None:
        reset_body_ids = torch.arange(J, device=state.device)

    progress_buf += 1

    # reward (:1230-1271)
    t = progress_buf * dt + motion_start_times + motion_start_times_offset
    ref = lib.get_motion_state(sampled_motion_ids, t, global_offset)
    reward, raw = imitation_reward(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang,
        ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"], rwd_specs,
    )  # fmt: skip

    power_reward = None
    if dof_force is not None:  # use_power_reward (:1297-1305)
        power = torch.abs(torch.multiply(dof_force, dof_vel)).sum(dim=-1)
        power_reward = -rew_power_coef * power
        power_reward[progress_buf <= 3] = 0
        reward = reward + power_reward
        raw = torch.cat([raw, power_reward[:, None]], dim=-1)  # reward_raw[:, -1] = power_reward

    # reset (:1313-1335); quirk: the reference compares against the un-indexed lengths,
    # valid because ids == arange(N); length[id] is the same thing there.
    pass_time = t >= lib._motion_lengths[sampled_motion_ids]
    reset, terminated = im_reset(
        reset_buf, progress_buf, None, None,
        pos[:, reset_body_ids].clone(), ref["rg_pos"][:, reset_body_ids].clone(),
        pass_time, enable_early_termination, termination_distances[reset_body_ids], use_mean,
    )  # fmt: skip

    # observations (:937-961)
    self_obs = self_obs_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    refs = []
    for k in range(1, time_steps + 1):
        tk = (progress_buf + k) * dt + motion_start_times + motion_start_times_offset
        refs.append(lib.get_motion_state(sampled_motion_ids, tk, global_offset))

    def stack(key):
        return torch.stack([r[key] for r in refs], dim=1).reshape((N * time_steps,) + refs[0][key].shape[1:])

    task_obs = imitation_obs_v6(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang,
        stack("rg_pos"), stack("rb_rot"), stack("body_vel"), stack("body_ang_vel"), time_steps, True,
    )  # fmt: skip
    obs = torch.cat([self_obs, task_obs], dim=-1)
    return obs, reward, raw, reset, terminated


# ----------------------------------------------------------------------------------
# actions -> PD targets — envs/humanoid_phc.py:105-128, 1218-1228
# ----------------------------------------------------------------------------------


def action_to_pd_targets(action, offset, scale, res_action=False, ref_dof_pos=None, dof_pos=None,
                         zero_joints=()):  # fmt: skip
    """_action_to_pd_targets (:1218-1228) followed by step()'s freeze_hand / freeze_toe zeroing
    (:118-127); ``zero_joints`` are indices into DOF_NAMES."""
    import numpy as np

    if res_action:
        pd = ref_dof_pos + scale * action
        pd = torch.maximum(torch.minimum(pd, dof_pos + np.pi / 2), dof_pos - np.pi / 2)
    else:
        pd = offset + scale * action
    for j in zero_joints:
        pd[:, 3 * j : 3 * j + 3] = 0
    return pd


# ----------------------------------------------------------------------------------
# reset — envs/humanoid_phc.py:665-778, motion_lib.py:526-535
# ----------------------------------------------------------------------------------


def sample_time_interval(motion_lengths: Tensor, motion_ids: Tensor, phase: Tensor) -> Tensor:
    """MotionLibBase.sample_time_interval, motion_lib.py:526-535, with the uniform numbers
    ``torch.rand`` would draw passed in as ``phase`` (truncate_time=None)."""
    motion_len = motion_lengths[motion_ids]
    curr_fps = 1 / 30
    return ((phase * motion_len) / curr_fps).long() * curr_fps


def reset_envs(
    lib: OracleMotionLib,
    env_ids: Tensor,
    phase: Tensor,  # [len(env_ids)]
    state: Tensor,  # [N, B, 13] AoS sim state — written in place
    root_states: Tensor,  # [N, 13]
    dof_pos: Tensor,  # [N, 69]
    dof_vel: Tensor,  # [N, 69]
    progress_buf: Tensor,
    reset_buf: Tensor,
    terminate_buf: Tensor,
    motion_start_times: Tensor,
    motion_start_times_offset: Tensor,
    global_offset: Tensor,
    sampled_motion_ids: Tensor,
    obs_buf: Tensor,
    dt: float,
    random_init: bool = True,
    flag_test: bool = False,
    time_steps: int = 1,
    num_bodies: int = 24,
    default_mask: Tensor = None,  # [len(env_ids)] bool: these envs take _reset_default (StateInit.Default: all of them;
    #                               StateInit.Hybrid: where torch.bernoulli(hybrid_init_prob) came up 0, :733-745)
    initial_root_states: Tensor = None,  # [N, 13]
    initial_dof_pos: Tensor = None,  # [N, 69]
    initial_dof_vel: Tensor = None,  # [N, 69]
):
    """HumanoidPHC._reset_envs(env_ids), minus the PhysX setters (humanoid_phc.py:665-676).  Reference-state init
    (StateInit.Random / Start, and the reference-init share of Hybrid): _sample_ref_state (:845-875) -> _set_env_state
    (:901-931) -> clock updates (:724-731); ``phase`` has one number per such env.  Default init (:688-692): root and
    dof state from the initial buffers — the rigid-body tensors and the motion clock stay as they are.  Then, for all
    of env_ids: buffer zeroing (:775-778) -> _compute_observations(env_ids) (:937-961).  All tensors are updated in
    place; returns the obs rows of env_ids."""
    J = num_bodies
    all_ids = env_ids
    if default_mask is not None:
        def_ids = env_ids[default_mask]
        env_ids = env_ids[~default_mask]
        # _reset_default
        root_states[def_ids] = initial_root_states[def_ids]
        dof_pos[def_ids] = initial_dof_pos[def_ids]
        dof_vel[def_ids] = initial_dof_vel[def_ids]
    if env_ids.shape[0] > 0:
        ids = sampled_motion_ids[env_ids]
        if random_init:
            motion_times = sample_time_interval(lib._motion_lengths, ids, phase)
        else:
            motion_times = torch.zeros(env_ids.shape[0])
        if flag_test:
            motion_times[:] = 0
        res = lib.get_motion_state(ids, motion_times, global_offset[env_ids])  # the old offset (:860)
        # _set_env_state
        root_states[env_ids, 0:3] = res["root_pos"]
        root_states[env_ids, 3:7] = res["root_rot"]
        root_states[env_ids, 7:10] = res["root_vel"]
        root_states[env_ids, 10:13] = res["root_ang_vel"]
        dof_pos[env_ids] = res["dof_pos"]
        dof_vel[env_ids] = res["dof_vel"]
        state[env_ids, :J, 0:3] = res["rg_pos"]
        state[env_ids, :J, 3:7] = res["rb_rot"]
        state[env_ids, :J, 7:10] = res["body_vel"]
        state[env_ids, :J, 10:13] = res["body_ang_vel"]
        # _reset_ref_state_init
        global_offset[env_ids] = 0
        motion_start_times[env_ids] = motion_times
        motion_start_times_offset[env_ids] = 0
    env_ids = all_ids
    ids = sampled_motion_ids[env_ids]
    # _reset_env_tensors
    progress_buf[env_ids] = 0
    reset_buf[env_ids] = 0
    terminate_buf[env_ids] = 0
    # _compute_observations(env_ids)
    pos, rot = state[env_ids, :J, 0:3], state[env_ids, :J, 3:7]
    vel, ang = state[env_ids, :J, 7:10], state[env_ids, :J, 10:13]
    self_obs = self_obs_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    refs = []
    for k in range(1, time_steps + 1):
        tk = (progress_buf[env_ids] + k) * dt + motion_start_times[env_ids] + motion_start_times_offset[env_ids]
        refs.append(lib.get_motion_state(ids, tk, global_offset[env_ids]))
    n = env_ids.shape[0]

    def stack(key):
        return torch.stack([r[key] for r in refs], dim=1).reshape((n * time_steps,) + refs[0][key].shape[1:])

    task = imitation_obs_v6(pos[:, 0], rot[:, 0], pos, rot, vel, ang, stack("rg_pos"), stack("rb_rot"),
                            stack("body_vel"), stack("body_ang_vel"), time_steps, True)  # fmt: skip
    obs = torch.cat([self_obs, task], dim=-1)
    obs_buf[env_ids] = obs
    return obs


# ----------------------------------------------------------------------------------
# observation normaliser — policies/running_norm.py
# ----------------------------------------------------------------------------------


def running_norm_update(running_mean, running_var, count, x):
    """RunningNorm.update, policies/running_norm.py:23-34.  Returns the new triple."""
    x = x.float()
    mean = x.mean(0, keepdim=True)
    var = x.var(0, unbiased=False, keepdim=True)
    weight = 1 / count
    return (
        running_mean * (1 - weight) + mean * weight,
        running_var * (1 - weight) + var * weight,
        count + 1,
    )


def running_norm_forward(running_mean, running_var, x, epsilon=1e-5, clip=10.0):
    """RunningNorm.forward, policies/running_norm.py:15-20."""
    return torch.clamp((x - running_mean.expand_as(x)) / torch.sqrt(running_var.expand_as(x) + epsilon), -clip, clip)


# ----------------------------------------------------------------------------------------
# episode bookkeeping of the pufferlib wrapper — clean_pufferl/env.py
# ----------------------------------------------------------------------------------------
def episode_update(state: Dict[str, Tensor], reset: Tensor, terminate: Tensor, rewards: Tensor, reward_raw: Tensor):
    """The per-env part of PHCPufferEnv.step, clean_pufferl/env.py:121-159, in place on ``state``.

    ``state`` holds ``terminals, truncations, masks`` (bool [n]), ``episode_returns`` (f32 [n]),
    ``episode_lengths`` (i32 [n]), ``raw_rewards`` (f32 [5]) and ``stats`` (f64 [4] = finished
    episodes, sum of their returns, sum of their lengths, how many were truncations: the running
    sums behind the three lists ``mean_and_log`` (:191-204) averages).  ``reset`` is ``reset_buf``
    as ``HumanoidPHC.step`` left it, i.e. before ``env.reset(reset_indices)`` clears it (:133-135,
    humanoid_phc.py:778); ``terminate`` is ``extras["terminate"]``.

    Reference order, kept: finished episodes are logged and zeroed first (:137-140); then, because
    ``env.reset`` has already cleared ``reset_buf``, ``~reset_buf`` is all-true at :158-159 and EVERY
    env — the just-reset ones too — accumulates this step's reward and one step of length."""
    reset = reset.bool()
    terminate = terminate.bool()
    state["raw_rewards"] += reward_raw.mean(dim=0)  # :124
    state["terminals"][:] = terminate  # :130,145-146 (terminated envs are always reset envs)
    trunc = reset & ~terminate  # :149-150
    state["truncations"][:] = trunc
    state["masks"][:] = ~trunc  # :132,154
    st = state["stats"]
    st[0] += reset.sum().double()
    st[1] += state["episode_returns"][reset].double().sum()
    st[2] += state["episode_lengths"][reset].double().sum()
    st[3] += trunc.sum().double()
    state["episode_returns"][reset] = 0  # :139-140
    state["episode_lengths"][reset] = 0
    state["episode_returns"] += rewards  # :158-159
    state["episode_lengths"] += 1
    return state


# ----------------------------------------------------------------------------------------
# AMP observation buffers of the env — envs/humanoid_phc.py:791-843, 1125-1212, 1341-1361
# ----------------------------------------------------------------------------------------
def amp_obs_from_state(state: Tensor, dof_pos: Tensor, dof_vel: Tensor, key_body_ids: Tensor, dof_subset: Tensor):
    """_compute_amp_observations -> _compute_amp_observations_from_state (:1125-1212) with the config's
    constant flags (local_root_obs, amp_root_height_obs, has_dof_subset, upright = True; shape / limb
    weight columns off).  ``state`` is the [n, B, 13] AoS rigid-body state; root = body 0."""
    return amp_obs_smpl(
        state[:, 0, 0:3], state[:, 0, 3:7], state[:, 0, 7:10], state[:, 0, 10:13], dof_pos, dof_vel,
        state[:, key_body_ids, 0:3], None, None, dof_subset, True, True, True, False, False, True,
    )  # fmt: skip


def update_hist_amp_obs(amp_obs_buf: Tensor, env_ids: Optional[Tensor] = None):
    """_update_hist_amp_obs (:1341-1350): slots 1.. take what slots 0..S-2 held."""
    if env_ids is None:
        amp_obs_buf[:, 1:] = amp_obs_buf[:, :-1].clone()
    else:
        amp_obs_buf[env_ids, 1:] = amp_obs_buf[env_ids, :-1].clone()


def init_amp_obs_ref(lib: OracleMotionLib, amp_obs_buf: Tensor, amp_obs_demo_buf: Tensor, env_ids: Tensor,
                     motion_ids: Tensor, motion_times: Tensor, dt: float, key_body_ids: Tensor, dof_subset: Tensor):  # fmt: skip
    """_init_amp_obs_ref (:805-819) with _get_amp_obs (:821-838): history slot k+1 of the reset envs is the
    AMP observation of the reference motion at ``motion_time - dt*(k+1)`` (no global offset), then the
    envs' whole rows are copied to the demo buffer."""
    S = amp_obs_buf.shape[1]
    ids = torch.tile(motion_ids.unsqueeze(-1), [1, S - 1]).view(-1)
    steps = -dt * (torch.arange(0, S - 1) + 1)
    times = (motion_times.unsqueeze(-1) + steps).view(-1)
    res = lib.get_motion_state(ids, times, None)
    obs = amp_obs_smpl(
        res["root_pos"], res["root_rot"], res["root_vel"], res["root_ang_vel"], res["dof_pos"], res["dof_vel"],
        res["rg_pos"][:, key_body_ids], None, None, dof_subset, True, True, True, False, False, True,
    )  # fmt: skip
    amp_obs_buf[env_ids, 1:] = obs.view(env_ids.shape[0], S - 1, -1)
    amp_obs_demo_buf[env_ids] = amp_obs_buf[env_ids]


class OracleEnv:
    """HumanoidPHC.step / reset (envs/humanoid_phc.py:90-172) with PhysX replaced by whatever the caller
    writes into ``state`` / ``dof_state`` / ``dof_force`` between the pre- and the post-physics half.
    Default config: power reward on, hands and toes frozen, reference-state init, T = 1."""

    def __init__(self, lib: OracleMotionLib, num_envs: int, progress_buf, motion_start_times,
                 motion_start_times_offset, global_offset, sampled_motion_ids, pd_action_offset, pd_action_scale,
                 dof_subset, key_body_ids, num_amp_obs_steps: int = 10, use_amp_obs: bool = True,
                 dt: float = 2 * (1.0 / 60.0), termination_distance: float = 0.25, rew_power_coef: float = 0.0005,
                 zero_joints=(17, 22, 3, 7)):  # fmt: skip
        N = num_envs
        self.lib, self.N, self.dt = lib, N, dt
        self.progress_buf = progress_buf.clone()
        self._motion_start_times = motion_start_times.clone()
        self._motion_start_times_offset = motion_start_times_offset.clone()
        self._global_offset = global_offset.clone()
        self._sampled_motion_ids = sampled_motion_ids.clone()
        self._pd_action_offset, self._pd_action_scale = pd_action_offset, pd_action_scale
        self.dof_subset, self._key_body_ids = dof_subset, key_body_ids
        self.use_amp_obs = use_amp_obs
        self.rew_power_coef = rew_power_coef
        self.flag_test = False  # toggle_eval_mode (:1427): resets start at time 0 (:852-853)
        self.zero_joints = zero_joints  # L_Hand, R_Hand, L_Toe, R_Toe in DOF_NAMES (:118-127)
        self._termination_distances = torch.full((24,), termination_distance)
        self.state = torch.zeros(N, 24, 13)
        self.root_states = torch.zeros(N, 13)
        self.dof_state = torch.zeros(N, 69, 2)
        self.dof_force = torch.zeros(N, 69)
        self.obs_buf = torch.zeros(N, 934)
        self.rew_buf = torch.zeros(N)
        self.reward_raw = torch.zeros(N, 5)
        self.reset_buf = torch.ones(N, dtype=torch.bool)
        self._terminate_buf = torch.ones(N, dtype=torch.bool)
        P = 13 + 23 * 6 + 69 + 3 * len(key_body_ids) - 9 * ((69 - len(dof_subset)) // 3)  # :470-476
        self._amp_obs_buf = torch.zeros(N, num_amp_obs_steps, P)
        self._amp_obs_demo_buf = torch.zeros_like(self._amp_obs_buf)

    def step(self, actions: Tensor, physics):
        pd_target = action_to_pd_targets(actions, self._pd_action_offset, self._pd_action_scale,
                                         zero_joints=self.zero_joints)  # fmt: skip
        physics(self)
        obs, rew, raw, reset, term = step(
            self.lib, self.state, self.progress_buf, self._motion_start_times, self._motion_start_times_offset,
            self._global_offset, self._sampled_motion_ids, self._termination_distances, self.dt,
            reset_buf=self.reset_buf, dof_force=self.dof_force, dof_vel=self.dof_state[..., 1],
            rew_power_coef=self.rew_power_coef,
        )  # fmt: skip
        self.obs_buf[:], self.rew_buf[:], self.reward_raw[:] = obs, rew, raw
        self.reset_buf[:], self._terminate_buf[:] = reset, term
        if self.use_amp_obs:  # :153-157
            update_hist_amp_obs(self._amp_obs_buf)
            self._amp_obs_buf[:, 0] = amp_obs_from_state(self.state, self.dof_state[..., 0], self.dof_state[..., 1],
                                                         self._key_body_ids, self.dof_subset)  # fmt: skip
        return pd_target

    def reset(self, env_ids: Tensor, phase: Tensor, default_mask: Tensor = None):
        """``default_mask`` [len(env_ids)]: the envs that take StateInit.Default's path (all of them in Default mode, the
        Bernoulli losers in Hybrid mode); ``phase`` then has one number per remaining env.  Needs ``initial_root_states``,
        ``initial_dof_pos``, ``initial_dof_vel`` set on the instance."""
        if len(env_ids) == 0:
            return
        kw = {}
        if default_mask is not None:
            assert not self.use_amp_obs, "_init_amp_obs raises NotImplementedError for default-reset envs (:795-797)"
            kw = dict(default_mask=default_mask, initial_root_states=self.initial_root_states,
                      initial_dof_pos=self.initial_dof_pos, initial_dof_vel=self.initial_dof_vel)  # fmt: skip
        reset_envs(
            self.lib, env_ids, phase, self.state, self.root_states, self.dof_state[..., 0], self.dof_state[..., 1],
            self.progress_buf, self.reset_buf, self._terminate_buf, self._motion_start_times,
            self._motion_start_times_offset, self._global_offset, self._sampled_motion_ids, self.obs_buf, self.dt,
            flag_test=self.flag_test, **kw,
        )  # fmt: skip
        if self.use_amp_obs:  # _init_amp_obs (:791-799)
            self._amp_obs_buf[env_ids, 0] = amp_obs_from_state(
                self.state[env_ids], self.dof_state[env_ids, :, 0], self.dof_state[env_ids, :, 1],
                self._key_body_ids, self.dof_subset)  # fmt: skip
            init_amp_obs_ref(self.lib, self._amp_obs_buf, self._amp_obs_demo_buf, env_ids,
                             self._sampled_motion_ids[env_ids], self._motion_start_times[env_ids], self.dt,
                             self._key_body_ids, self.dof_subset)  # fmt: skip

    def resample_motions(self, new_lib: OracleMotionLib, phase: Tensor):
        """resample_motions, humanoid_phc.py:1363-1379 (training branch).  ``new_lib`` is what ``load_motions``
        built (oracle/build_oracle.py); get_root_pos_smpl (motion_lib.py:628-653) is the root row of the same
        position blend ``get_motion_state`` does, without the offset."""
        self.lib = new_lib
        time = self.progress_buf * self.dt + self._motion_start_times + self._motion_start_times_offset
        root_pos = new_lib.get_motion_state(self._sampled_motion_ids, time)["root_pos"]
        self._global_offset[:, :2] = self.root_states[:, :2] - root_pos[:, :2]
        self.reset(torch.arange(self.N), phase)
