"""CPU oracle for the motion-library build.  TEST INFRASTRUCTURE ONLY.

A restatement of what ``MotionLibSMPL.load_motions`` computes per clip
(PHC/motion_lib.py:257-428 and its worker ``load_motion_with_skeleton`` :748-824): the optional
random heading, forward kinematics of the SMPL tree, finite-difference + gaussian-filtered body
velocities, angular velocities from consecutive global rotations, dof velocities from consecutive
local rotations, and the per-clip metadata.  It issues the same numpy / scipy / ATen CPU ops in the
same order and **in the same dtypes** as the reference — that matters, because the reference mixes
them (all paths relative to /root/reference/packages/puffer-phc/puffer_phc/):

* ``local_rotation`` is computed in fp64 from the fp64 global rotations but stored into an fp32
  tensor (``quat_identity_like`` builds fp32, poselib_skeleton.py:575-594, torch_utils.py:200-216);
* the root translation is stored into the tree's fp32 ``local_translation`` (:606-620), so the whole
  forward-kinematics chain (:519-539, ``transform_mul`` torch_utils.py:322-330) runs in fp32;
* ``np.gradient`` therefore differentiates fp32 positions in fp32, scipy's ``gaussian_filter1d``
  accumulates in fp64 and rounds back to fp32 (:1231-1238);
* angular velocities stay fp64 until the final ``.float()`` (:1241-1251, motion_lib.py:401);
* ``compute_motion_dof_vels_jit`` (motion_lib.py:120-142) works on the fp32 local rotations.

Parity pin: ``tests/golden/motion_build.npz`` is produced by the reference's own ``load_motions``
(``tests/golden/make_golden.py motion_build``); ``tests/test_oracle_golden.py`` checks this module
against it.  Only ``tests/`` may import this module.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d
from scipy.spatial.transform import Rotation as sRot

from .phc_oracle import angle_axis, quat_conj, quat_mul

Tensor = torch.Tensor


def quat_normalize(q: Tensor) -> Tensor:
    """quat_unit(quat_pos(q)), torch_utils.py:154-196.  ``z`` is fp32 whatever ``q`` is (``.float()``),
    so an fp64 ``q`` stays fp64 by promotion."""
    z = (q[..., 3:] < 0).float()
    q = (1 - 2 * z) * q
    return q / q.norm(p=2, dim=-1).unsqueeze(-1).clamp(min=1e-9)


def quat_mul_norm(a: Tensor, b: Tensor) -> Tensor:
    """torch_utils.py:254-259."""
    return quat_normalize(quat_mul(a, b))


def quat_rotate(rot: Tensor, vec: Tensor) -> Tensor:
    """torch_utils.py:263-268: imaginary part of rot * (vec, 0) * conj(rot), two 8-multiply products."""
    other = torch.cat([vec, torch.zeros_like(vec[..., :1])], dim=-1)
    return quat_mul(quat_mul(rot, other), quat_conj(rot))[..., :3]


def quat_angle_axis(x: Tensor):
    """poselib's variant, torch_utils.py:219-228: angle = acos(clamp(2w^2-1)), axis = xyz / max(|xyz|, 1e-9)."""
    s = 2 * (x[..., 3] ** 2) - 1
    angle = s.clamp(-1, 1).arccos()
    axis = x[..., :3]
    axis = axis / axis.norm(p=2, dim=-1, keepdim=True).clamp(min=1e-9)
    return angle, axis


def local_rotation(global_rot: Tensor, parents: Sequence[int]) -> Tensor:
    """SkeletonState.local_rotation for is_local=False, poselib_skeleton.py:575-594 (fp32 result)."""
    out = torch.zeros(global_rot.shape, dtype=torch.float32)
    for j, p in enumerate(parents):
        if p == -1:
            out[..., j, :] = global_rot[..., j, :]
        else:
            out[..., j, :] = quat_mul_norm(quat_conj(global_rot[..., p, :]), global_rot[..., j, :])
    return out


def forward_kinematics(local_rot: Tensor, root_trans: Tensor, tree_local_translation: Tensor, parents: Sequence[int]):
    """global_transformation, poselib_skeleton.py:519-539, on cat(local_rotation, local_translation)
    (:596-620); everything is fp32 because both halves are."""
    B = local_rot.shape[0]
    lt = tree_local_translation.broadcast_to(B, *tree_local_translation.shape).clone()
    lt[..., 0, :] = root_trans  # fp64 -> the tree's dtype (fp32)
    g_rot, g_pos = [], []
    for j, p in enumerate(parents):
        if p == -1:
            g_rot.append(local_rot[..., j, :])
            g_pos.append(lt[..., j, :])
        else:  # transform_mul, torch_utils.py:322-330
            g_rot.append(quat_mul_norm(g_rot[p], local_rot[..., j, :]))
            g_pos.append(quat_rotate(g_rot[p], lt[..., j, :]) + g_pos[p])
    return torch.stack(g_pos, dim=-2)


def compute_velocity(p: Tensor, time_delta: float) -> Tensor:
    """SkeletonMotion._compute_velocity, poselib_skeleton.py:1231-1238."""
    velocity = np.gradient(p.numpy(), axis=-3) / time_delta
    return torch.from_numpy(gaussian_filter1d(velocity, 2, axis=-3, mode="nearest")).to(p)


def compute_angular_velocity(r: Tensor, time_delta: float) -> Tensor:
    """SkeletonMotion._compute_angular_velocity, poselib_skeleton.py:1241-1251 (fp64 in, fp64 out)."""
    diff = torch.zeros_like(r)
    diff[..., 3] = 1
    diff[..., :-1, :, :] = quat_mul_norm(r[..., 1:, :, :], quat_conj(r[..., :-1, :, :]))
    angle, axis = quat_angle_axis(diff)
    av = axis * angle.unsqueeze(-1) / time_delta
    return torch.from_numpy(gaussian_filter1d(av.numpy(), 2, axis=-3, mode="nearest"))


def dof_vels(local_rot: Tensor, fps: int) -> Tensor:
    """compute_motion_dof_vels_jit, motion_lib.py:120-142.  One frame pair at a time like the reference:
    ATen's vectorised loops send the tail of a 24-element row through libm and the rest through Sleef,
    so batching the frames would change some results by one ulp."""
    dt = 1.0 / fps
    rows = []
    for f in range(local_rot.shape[0] - 1):
        diff = quat_mul(quat_conj(local_rot[f]), local_rot[f + 1])
        angle, axis = angle_axis(diff)
        rows.append((axis * angle.unsqueeze(-1) / dt)[1:, :].flatten())
    rows.append(rows[-1])
    return torch.stack(rows, dim=0).view(local_rot.shape[0], -1, 3)


def random_heading(pose_aa: np.ndarray, pose_quat_global: np.ndarray, trans: Tensor, u: float):
    """The heading randomisation of load_motion_with_skeleton, motion_lib.py:789-799, with the uniform
    number ``u`` supplied (the reference draws ``np.random.random()``).  ``pose_aa`` is modified IN PLACE
    like the reference's (``to_torch`` of a numpy slice shares the file's memory, :783/:794)."""
    B, J, N = pose_quat_global.shape
    random_rot = np.zeros(3)
    random_rot[2] = np.pi * (2 * u - 1.0)
    h = sRot.from_euler("xyz", random_rot)
    pose_aa[:, :3] = (h * sRot.from_rotvec(pose_aa[:, :3])).as_rotvec()
    pose_quat_global = (h * sRot.from_quat(pose_quat_global.reshape(-1, 4))).as_quat().reshape(B, J, N)
    trans = torch.matmul(trans, torch.from_numpy(h.as_matrix().T))
    return pose_quat_global, trans


def build_motion_library(
    pose_quat_global: np.ndarray, root_trans: np.ndarray, pose_aa: np.ndarray, num_frames: Sequence[int],
    fps: Sequence[int], parents: Sequence[int], local_translation: np.ndarray, gender_betas: np.ndarray,
    limb_weights: np.ndarray, heading_u: Optional[Sequence[float]] = None, max_length: int = -1,
    crop_start: Optional[Sequence[int]] = None,
) -> Dict[str, Tensor]:  # fmt: skip
    """The files' clips are concatenated on the frame axis (``num_frames`` = the files' lengths);
    ``local_translation`` is [M,24,3] fp32 (one tree per clip).  ``max_length`` / ``crop_start`` are the
    window of :773-778 (the reference draws the start with ``random.randint``).  Returns the reference's
    attribute set (motion_lib.py:390-414) without the leading underscore.

    Reference quirk kept: ``_motion_aa`` is built from the FILE's ``pose_aa`` (:377), i.e. the uncropped
    clip — with the rows of the window heading-rotated in place (:783, :794) — so with a crop it has more
    rows than ``gts`` and is not frame-aligned with it."""
    parents = [int(p) for p in parents]
    cols = {k: [] for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs", "motion_aa")}
    lengths, dts, fpss, nfs = [], [], [], []
    file_start = 0
    for m, seq_len in enumerate(num_frames):
        seq_len, f = int(seq_len), int(fps[m])
        if max_length == -1 or seq_len < max_length:
            start, end = 0, seq_len
        else:
            start = int(crop_start[m])
            end = start + max_length
        file_aa = np.array(pose_aa[file_start : file_start + seq_len], dtype=np.float64)  # this clip's "file"
        sl = slice(file_start + start, file_start + end)
        file_start += seq_len
        nf = end - start
        q = np.asarray(pose_quat_global[sl], dtype=np.float64)
        t = torch.from_numpy(np.asarray(root_trans[sl], dtype=np.float64))
        if heading_u is not None:
            q, t = random_heading(file_aa[start:end], q, t, float(heading_u[m]))
        q = torch.from_numpy(q)
        lr = local_rotation(q, parents)
        gt = forward_kinematics(lr, t, torch.from_numpy(np.asarray(local_translation[m], dtype=np.float32)), parents)
        cols["gts"].append(gt)
        cols["grs"].append(q.float())
        cols["lrs"].append(lr)
        cols["gvs"].append(compute_velocity(gt, 1 / f).float())
        cols["gavs"].append(compute_angular_velocity(q, 1 / f).float())
        cols["dvs"].append(dof_vels(lr, f))
        cols["motion_aa"].append(torch.from_numpy(file_aa).float())  # :377, :391
        fpss.append(f)
        dts.append(1.0 / f)
        lengths.append(1.0 / f * (nf - 1))  # :372
        nfs.append(nf)
    out = {k: torch.cat(v, dim=0) for k, v in cols.items()}
    out["grvs"] = out["gvs"][:, 0]
    out["gravs"] = out["gavs"][:, 0]
    nfr = torch.tensor(nfs)
    shifted = nfr.roll(1)
    shifted[0] = 0
    out["length_starts"] = shifted.cumsum(0)  # :405-408
    out["motion_num_frames"] = nfr
    out["motion_lengths"] = torch.tensor(lengths, dtype=torch.float32)
    out["motion_dt"] = torch.tensor(dts, dtype=torch.float32)
    out["motion_fps"] = torch.tensor(fpss, dtype=torch.float32)
    out["motion_bodies"] = torch.from_numpy(np.asarray(gender_betas)).float()
    out["motion_limb_weights"] = torch.from_numpy(np.asarray(limb_weights)).float()
    return out
