"""Drop-ins for the obs / reward / reset functions of PHC/envs/common.py.

Same names, argument order, tensor layouts and return values as the reference's
``compute_humanoid_observations_smpl_max`` (:23), ``compute_imitation_observations_v6``
(:107), ``compute_imitation_reward`` (:271) and ``compute_humanoid_im_reset`` (:326); each
is one CUDA kernel launch through the C ABI.  Inputs may be the stride-13 views of the AoS
sim tensor (humanoid_phc.py:546-549) or contiguous tensors.  ``compute_imitation_
observations_v7`` does not exist in the reference; it is the position/velocity column
subset of v6.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _cabi


def compute_humanoid_observations_smpl_max(
    body_pos, body_rot, body_vel, body_ang_vel, smpl_params, limb_weight_params,
    local_root_obs, root_height_obs, upright, has_smpl_params, has_limb_weight_params,
):  # fmt: skip
    body, keep = _cabi.body_state(body_pos, body_rot, body_vel, body_ang_vel)
    n, J = keep[0].shape[0], keep[0].shape[1]
    dev = keep[0].device
    width = (1 if root_height_obs else 0) + 15 * J - 3
    extra = []
    if has_smpl_params:
        extra.append(smpl_params)
    if has_limb_weight_params:
        extra.append(limb_weight_params)
    total = width + sum(int(x.shape[-1]) for x in extra)
    out = torch.empty((n, total), dtype=torch.float32, device=dev)
    flags = (
        (_cabi.OBS_LOCAL_ROOT if local_root_obs else 0)
        | (_cabi.OBS_ROOT_HEIGHT if root_height_obs else 0)
        | (_cabi.OBS_UPRIGHT if upright else 0)
    )
    _cabi.check(
        _cabi.load().phc_self_obs_smpl_max(C.byref(body), n, flags, out.data_ptr(), out.stride(0), _cabi.stream_ptr(dev)),
        "phc_self_obs_smpl_max",
    )
    col = width
    for x in extra:  # appended verbatim (common.py:96-100): a device copy, no arithmetic
        out[:, col : col + x.shape[-1]] = x
        col += x.shape[-1]
    return out


def _imitation_obs(mode, root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
                   ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps, upright):  # fmt: skip
    body, keep = _cabi.body_state(body_pos, body_rot, body_vel, body_ang_vel)
    B, J = keep[0].shape[0], keep[0].shape[1]
    T = int(time_steps)
    dev = keep[0].device

    def as_ref(x, c):
        return x.reshape(B * T, J, c)

    ref, keep_ref = _cabi.body_state(as_ref(ref_body_pos, 3), as_ref(ref_body_rot, 4), as_ref(ref_body_vel, 3),
                                     as_ref(ref_body_ang_vel, 3), prefix="ref_")  # fmt: skip
    _cabi.require_cuda(root_pos, "root_pos", torch.float32)
    _cabi.require_cuda(root_rot, "root_rot", torch.float32)
    if root_pos.stride(-1) != 1:
        root_pos = root_pos.contiguous()
    if root_rot.stride(-1) != 1:
        root_rot = root_rot.contiguous()
    per = (24 if mode == 6 else 9) * J
    out = torch.empty((B, per * T), dtype=torch.float32, device=dev)
    _cabi.check(
        _cabi.load().phc_imitation_obs(
            root_pos.data_ptr(), root_pos.stride(0), root_rot.data_ptr(), root_rot.stride(0),
            C.byref(body), C.byref(ref), B, T, 1 if upright else 0, mode, out.data_ptr(), out.stride(0),
            _cabi.stream_ptr(dev),
        ),
        "phc_imitation_obs",
    )  # fmt: skip
    return out


def compute_imitation_observations_v6(
    root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
    ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps, upright,
):  # fmt: skip
    return _imitation_obs(6, root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
                          ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps, upright)  # fmt: skip


def compute_imitation_observations_v7(
    root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
    ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps, upright,
):  # fmt: skip
    return _imitation_obs(7, root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
                          ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, time_steps, upright)  # fmt: skip


def dof_subset_smpl(device=None) -> torch.Tensor:
    """The 57-dof subset the AMP observation uses: all joints but toes and hands
    (humanoid_phc.py:186-194 with body_sets.py:42)."""
    removed = (3, 7, 17, 22)  # L_Toe, R_Toe, L_Hand, R_Hand in DOF_NAMES
    idx = [3 * j + k for j in range(23) if j not in removed for k in range(3)]
    return torch.tensor(idx, dtype=torch.long, device=device)


def build_amp_observations_smpl(
    root_pos, root_rot, root_vel, root_ang_vel, dof_pos, dof_vel, key_body_pos, shape_params, limb_weight_params,
    dof_subset, local_root_obs, root_height_obs, has_dof_subset, has_shape_obs_disc, has_limb_weight_obs, upright,
):  # fmt: skip
    """Drop-in for envs/common.py:192-267 (one kernel launch)."""
    for t, nm in ((root_pos, "root_pos"), (root_rot, "root_rot"), (root_vel, "root_vel"),
                  (root_ang_vel, "root_ang_vel"), (dof_pos, "dof_pos"), (dof_vel, "dof_vel")):  # fmt: skip
        _cabi.require_cuda(t, nm, torch.float32)
    B, D = dof_pos.shape
    dev = root_pos.device

    def rows(t):
        return t if t.stride(-1) == 1 else t.contiguous()

    root_pos, root_rot, root_vel, root_ang_vel = rows(root_pos), rows(root_rot), rows(root_vel), rows(root_ang_vel)
    vk, key_body_pos = _cabi.view3(key_body_pos, "key_body_pos")
    K = key_body_pos.shape[1]
    sub = None
    num_sel = D
    if has_dof_subset:
        _cabi.require_cuda(dof_subset, "dof_subset", torch.int64)
        sub = dof_subset.contiguous()
        num_sel = sub.numel()
    a = _cabi.PhcAmpArgs(
        root_pos.data_ptr(), root_pos.stride(0), root_rot.data_ptr(), root_rot.stride(0),
        root_vel.data_ptr(), root_vel.stride(0), root_ang_vel.data_ptr(), root_ang_vel.stride(0),
        dof_pos.data_ptr(), dof_pos.stride(0), dof_pos.stride(1), dof_vel.data_ptr(), dof_vel.stride(0), dof_vel.stride(1),
        vk, K, _cabi.ptr(sub), num_sel,
        (_cabi.OBS_LOCAL_ROOT if local_root_obs else 0) | (_cabi.OBS_ROOT_HEIGHT if root_height_obs else 0)
        | (_cabi.OBS_UPRIGHT if upright else 0),
    )  # fmt: skip
    width = (1 if root_height_obs else 0) + 12 + 3 * num_sel + 3 * K
    extra = []
    if has_shape_obs_disc:
        extra.append(shape_params)
    if has_limb_weight_obs:
        extra.append(limb_weight_params)
    out = torch.empty((B, width + sum(int(x.shape[-1]) for x in extra)), dtype=torch.float32, device=dev)
    _cabi.check(
        _cabi.load().phc_amp_obs(C.byref(a), B, out.data_ptr(), out.stride(0), _cabi.stream_ptr(dev)), "phc_amp_obs"
    )
    col = width
    for x in extra:  # appended verbatim (common.py:261-264)
        out[:, col : col + x.shape[-1]] = x
        col += x.shape[-1]
    return out


def compute_imitation_reward(
    root_pos, root_rot, body_pos, body_rot, body_vel, body_ang_vel,
    ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, rwd_specs: Dict[str, float],
) -> Tuple[torch.Tensor, torch.Tensor]:  # fmt: skip
    # root_pos / root_rot are accepted and, as in the reference (common.py:272-273), unused
    body, keep = _cabi.body_state(body_pos, body_rot, body_vel, body_ang_vel)
    ref, keep_ref = _cabi.body_state(ref_body_pos, ref_body_rot, ref_body_vel, ref_body_ang_vel, prefix="ref_")
    n = keep[0].shape[0]
    dev = keep[0].device
    spec = _cabi.reward_spec(rwd_specs)
    reward = torch.empty(n, dtype=torch.float32, device=dev)
    raw = torch.empty((n, 4), dtype=torch.float32, device=dev)
    _cabi.check(
        _cabi.load().phc_imitation_reward(
            C.byref(body), C.byref(ref), n, C.byref(spec), reward.data_ptr(), raw.data_ptr(), raw.stride(0),
            _cabi.stream_ptr(dev),
        ),
        "phc_imitation_reward",
    )
    return reward, raw


def compute_humanoid_im_reset(
    reset_buf, progress_buf, contact_buf, contact_body_ids, rigid_body_pos, ref_body_pos, pass_time,
    enable_early_termination, termination_distance, use_mean,
):  # fmt: skip
    # contact_buf / contact_body_ids are never read by the reference (common.py:329-330)
    vp, rigid_body_pos = _cabi.view3(rigid_body_pos, "rigid_body_pos")
    vr, ref_body_pos = _cabi.view3(ref_body_pos, "ref_body_pos")
    n, R = rigid_body_pos.shape[0], rigid_body_pos.shape[1]
    dev = rigid_body_pos.device
    _cabi.require_cuda(progress_buf, "progress_buf", torch.int16)
    _cabi.require_cuda(pass_time, "pass_time", torch.bool)
    _cabi.require_cuda(termination_distance, "termination_distance", torch.float32)
    progress_buf = progress_buf.contiguous()
    pass_time = pass_time.contiguous()
    termination_distance = termination_distance.contiguous()
    if termination_distance.numel() < (1 if use_mean else R):
        raise _cabi.PhcError("termination_distance has fewer entries than reset bodies")
    reset = torch.empty(n, dtype=torch.bool, device=dev)
    terminated = torch.empty(n, dtype=torch.bool, device=dev)
    _cabi.check(
        _cabi.load().phc_im_reset(
            C.byref(vp), C.byref(vr), R, progress_buf.data_ptr(), pass_time.data_ptr(),
            termination_distance.data_ptr(), 1 if enable_early_termination else 0, 1 if use_mean else 0, n,
            reset.data_ptr(), terminated.data_ptr(), _cabi.stream_ptr(dev),
        ),
        "phc_im_reset",
    )  # fmt: skip
    return reset.to(reset_buf.dtype), terminated.to(reset_buf.dtype)
