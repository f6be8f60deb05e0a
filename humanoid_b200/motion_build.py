"""Motion-library build on the device: ``MotionLibSMPL.load_motions`` with the reference's signature.

Drop-in for the load side of the reference's ``MotionLibSMPL`` (PHC/motion_lib.py:257-428 and the
per-clip worker ``load_motion_with_skeleton`` :748-824).  The reference spends ~20 s per 4096 clips in
~1000 small ATen / numpy / scipy calls per clip spread over worker processes; here the clips' fp64
arrays are concatenated on the frame axis in pinned host memory, copied once, and three kernels
(``phc_motion_build``) produce ``gts, grs, lrs, gvs, gavs, dvs, _motion_aa`` for all frames at once,
already HBM-resident where ``get_motion_state`` and the fused step read them.

The host side keeps what is host logic in the reference: which clips are sampled (:300-310), the
``max_length`` crop (:773-778), the random heading draw (:789-791, one ``np.random.random()`` per
clip, in clip order, so a seeded run draws the same headings) and the per-clip metadata (:361-392).
``fix_trans_height`` (:696-745) needs the SMPL mesh model files and is not implemented — this is the
``mesh_parsers is None`` path (:692-694, :801).  Two quirks of the reference's single-process loader
(``num_thread == 1``) are kept rather than fixed: ``_motion_aa`` is the files' UNCROPPED ``pose_aa`` (:377), so
with a ``max_length`` crop it has more rows than ``gts``; and the random heading is written in place into the
caller's ``pose_aa`` window (:783, :794), so it accumulates over resamples.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _cabi
from .motion_lib import MotionLib

FILTER_SIGMA = 2.0  # SkeletonMotion._compute_velocity / _compute_angular_velocity, poselib_skeleton.py:1236,1250


@dataclass
class SkeletonTree:
    """The three fields of poselib's SkeletonTree (poselib_skeleton.py:96-118) the build reads."""

    node_names: Sequence[str]
    parent_indices: torch.Tensor  # [J] int, -1 for the root
    local_translation: torch.Tensor  # [J,3] fp32


def gaussian_taps(sigma: float = FILTER_SIGMA, radius: int = _cabi.BUILD_FILTER_RADIUS) -> np.ndarray:
    """scipy.ndimage's ``_gaussian_kernel1d(sigma, 0, radius)``: exp(-x^2 / (2 sigma^2)), normalised (fp64)."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    return phi / phi.sum()


def heading_half_angle(u: np.ndarray) -> np.ndarray:
    """(sin, cos) of half the heading angle pi*(2u-1) (motion_lib.py:790): the z and w of
    ``sRot.from_euler("xyz", [0, 0, angle])``."""
    half = np.pi * (2 * np.asarray(u, dtype=np.float64) - 1.0) / 2
    return np.stack([np.sin(half), np.cos(half)], axis=-1)


def heading_on_rotvec(v: np.ndarray, z: float, w: float) -> np.ndarray:
    """``(heading * Rotation.from_rotvec(v)).as_rotvec()`` for the heading quaternion (0, 0, z, w), in scipy's
    formulas (Taylor branches below 1e-3 rad), fp64, vectorised over rows: the root entry of ``pose_aa``
    under the random heading (motion_lib.py:794)."""
    v = np.asarray(v, dtype=np.float64)
    ang = np.sqrt((v * v).sum(-1))
    small = ang <= 1e-3
    a2 = ang * ang
    safe = np.where(small, 1.0, ang)
    sc = np.where(small, 0.5 - a2 / 48.0 + a2 * a2 / 3840.0, np.sin(safe / 2) / safe)
    x, y, zq, wq = sc * v[..., 0], sc * v[..., 1], sc * v[..., 2], np.cos(ang / 2)
    x, y, zq, wq = w * x - z * y, w * y + z * x, w * zq + z * wq, w * wq - z * zq
    sign = np.where(wq < 0, -1.0, 1.0)
    x, y, zq, wq = sign * x, sign * y, sign * zq, sign * wq
    a = 2 * np.arctan2(np.sqrt(x * x + y * y + zq * zq), wq)
    small = a <= 1e-3
    a2 = a * a
    s2 = np.where(small, 2 + a2 / 12 + 7 * a2 * a2 / 2880, a / np.sin(np.where(small, 1.0, a) / 2))
    return np.stack([s2 * x, s2 * y, s2 * zq], axis=-1)


def _as_np64(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def build_motion_tensors(
    pose_quat_global, root_trans, pose_aa, num_frames, fps, parent_indices, local_translation,
    heading_u=None, device="cuda", pin: bool = False,
) -> Dict[str, torch.Tensor]:  # fmt: skip
    """Clips concatenated on the frame axis -> the device tensors of motion_lib.py:396-403.

    ``pose_quat_global [F,24,4]``, ``root_trans [F,3]``, ``pose_aa [F,72]`` (or None) are fp64 as in the pkl
    (scripts/phc_convert_amass_data.py:186-194); ``local_translation [M,24,3]`` fp32 is one skeleton per clip;
    ``heading_u [M]`` are the uniform numbers of the random heading, None for the deterministic path.
    ``pin`` stages the inputs through freshly pinned memory first; for a one-shot upload the pageable copy is
    faster (allocating 1 GB of pinned memory costs more than it saves, ``profiles/r1_motion_build.md``)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _cabi.PhcError("the motion-library build runs on the device: pass device='cuda' (there is no CPU path)")
    nf = np.asarray(num_frames, dtype=np.int64)
    M, F = int(nf.shape[0]), int(nf.sum())
    starts = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(np.int64) if M else np.zeros(0, np.int64)
    quat, trans = _as_np64(pose_quat_global), _as_np64(root_trans)
    if quat.shape != (F, _cabi.NUM_BODIES, 4) or trans.shape != (F, 3):
        raise _cabi.PhcError(f"expected pose_quat_global [{F},24,4] and root_trans [{F},3], got {quat.shape}, {trans.shape}")
    lt = np.asarray(_as_np64(local_translation), dtype=np.float32)
    if lt.shape != (M, _cabi.NUM_BODIES, 3):
        raise _cabi.PhcError(f"local_translation must be [{M},24,3], got {lt.shape}")
    if M and int(nf.min()) < 1:
        raise _cabi.PhcError("every clip needs at least one frame")

    def up(a, dtype):  # one copy each
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        return t.pin_memory().to(dev, non_blocking=True) if pin and t.numel() else t.to(dev)

    d_quat, d_trans = up(quat, torch.float64), up(trans, torch.float64)
    d_aa = None
    if pose_aa is not None:
        aa = _as_np64(pose_aa).reshape(F, -1)
        if aa.shape[1] != _cabi.NUM_BODIES * 3:
            raise _cabi.PhcError(f"pose_aa must be [{F},72], got {aa.shape}")
        d_aa = up(aa, torch.float64)
    d_lt, d_nf, d_starts = up(lt, torch.float32), up(nf, torch.int64), up(starts, torch.int64)
    d_fps = up(np.asarray(fps, dtype=np.float64), torch.float64)
    d_head = up(heading_half_angle(heading_u), torch.float64) if heading_u is not None else None

    def f32(*shape):
        return torch.empty(shape, dtype=torch.float32, device=dev)

    J = _cabi.NUM_BODIES
    out = {"gts": f32(F, J, 3), "grs": f32(F, J, 4), "lrs": f32(F, J, 4), "gvs": f32(F, J, 3), "gavs": f32(F, J, 3),
           "dvs": f32(F, J - 1, 3), "motion_aa": f32(F, J * 3)}  # fmt: skip
    scratch = torch.empty(F * 108, dtype=torch.float64, device=dev)  # PhcBuildArgs.scratch
    parents = (C.c_int32 * J)(*[int(p) for p in np.asarray(parent_indices).tolist()])
    taps = (C.c_double * (2 * _cabi.BUILD_FILTER_RADIUS + 1))(*gaussian_taps().tolist())
    if d_aa is None:
        out["motion_aa"].zero_()  # clips without "beta": zeros (motion_lib.py:380)
    args = _cabi.PhcBuildArgs(
        d_quat.data_ptr(), d_trans.data_ptr(), _cabi.ptr(d_aa), d_lt.data_ptr(), d_nf.data_ptr(), d_starts.data_ptr(),
        d_fps.data_ptr(), _cabi.ptr(d_head), parents, taps, F, M,
        out["gts"].data_ptr(), out["grs"].data_ptr(), out["lrs"].data_ptr(), out["gvs"].data_ptr(),
        out["gavs"].data_ptr(), out["dvs"].data_ptr(), out["motion_aa"].data_ptr() if d_aa is not None else None,
        scratch.data_ptr(),
    )  # fmt: skip
    _cabi.check(_cabi.load().phc_motion_build(C.byref(args), _cabi.stream_ptr(dev)), "phc_motion_build")
    out["length_starts"] = d_starts
    out["motion_num_frames"] = d_nf
    return out


class ClipSampler:
    """Which clips a load draws: the host-side state and methods of ``MotionLibBase`` that involve no frame data
    (``load_data`` :192-230, ``setup_constants`` :232-245, the PMCP sampling weights :454-500, ``sample_motions``
    :510-513).  Pure host logic, no device needed."""

    def _init_clips(self, motion_data, min_length: int = -1, im_eval: bool = False):
        if isinstance(motion_data, (str, bytes)) or hasattr(motion_data, "__fspath__"):
            import joblib  # a pkl of {clip name: entry}, scripts/phc_convert_amass_data.py:186-194

            motion_data = joblib.load(motion_data)
        if min_length != -1:  # :205-208
            motion_data = {k: v for k, v in motion_data.items() if len(v["pose_quat_global"]) >= min_length}
        elif im_eval:  # longest first (stable), :209-217
            motion_data = dict(sorted(motion_data.items(), key=lambda e: len(e[1]["pose_quat_global"]), reverse=True))
        self._motion_data_keys = np.array(list(motion_data.keys()))
        self._motion_data_list = list(motion_data.values())
        n = self._num_unique_motions = len(self._motion_data_list)
        self._curr_motion_ids = None
        self._termination_history = torch.zeros(n)
        self._success_rate = torch.zeros(n)
        self._sampling_history = torch.zeros(n)
        self._sampling_prob = torch.ones(n) / max(n, 1)
        self._sampling_batch_prob = None

    def pick_clips(self, n: int, random_sample: bool, start_idx: int, sample_idxes, is_deterministic: bool):
        """load_motions :300-321."""
        if sample_idxes is None or len(sample_idxes) != n:
            if not is_deterministic and random_sample:
                sample_idxes = torch.multinomial(self._sampling_prob, num_samples=n, replacement=True)
            else:
                sample_idxes = torch.remainder(torch.arange(n) + start_idx, self._num_unique_motions)
        sample_idxes = torch.as_tensor(sample_idxes).cpu()
        self.curr_motion_keys = self._motion_data_keys[sample_idxes.numpy()]
        picked = self._sampling_prob[sample_idxes]
        self._sampling_batch_prob = picked / picked.sum()
        return sample_idxes

    def update_hard_sampling_weight(self, failed_keys):  # :454-470
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._sampling_prob[:] = 0
            self._sampling_prob[indexes] = 1 / len(indexes)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions) / self._num_unique_motions

    def update_soft_sampling_weight(self, failed_keys):  # :472-492
        if len(failed_keys) > 0:
            all_keys = self._motion_data_keys.tolist()
            indexes = [all_keys.index(k) for k in failed_keys]
            self._termination_history[indexes] += 1
            self.update_sampling_prob(self._termination_history)
        else:
            self._sampling_prob = torch.ones(self._num_unique_motions) / self._num_unique_motions

    def update_sampling_prob(self, termination_history):  # :494-500
        if len(termination_history) == len(self._termination_history) and termination_history.sum() > 0:
            self._sampling_prob[:] = termination_history / termination_history.sum()
            self._termination_history = termination_history
            return True
        return False

    def sample_motions(self, n):  # :510-513
        return torch.multinomial(self._sampling_batch_prob, num_samples=n, replacement=True)


class MotionLibSMPL(ClipSampler, MotionLib):
    """``MotionLibSMPL(motion_data, device, ...)`` then ``load_motions(skeleton_trees, gender_betas,
    limb_weights, ...)`` as in the reference (motion_lib.py:676-694, :257); queries are inherited."""

    def __init__(self, motion_data, device="cuda", max_length: int = -1, is_deterministic: bool = False,
                 im_eval: bool = False, min_length: int = -1, step_dt: float = 1 / 30):  # fmt: skip
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _cabi.PhcError("MotionLib lives in HBM: pass device='cuda' (there is no CPU path)")
        self._device = dev
        self._handle = None
        self.max_length, self.is_deterministic, self.im_eval = int(max_length), bool(is_deterministic), bool(im_eval)
        self._sim_fps = 1 / step_dt  # :183
        self._init_clips(motion_data, min_length=min_length, im_eval=im_eval)

    def load_motions(self, skeleton_trees, gender_betas, limb_weights, random_sample=True, start_idx=0, max_len=-1,
                     sample_idxes=None) -> List[dict]:  # fmt: skip
        sample_idxes = self.pick_clips(len(skeleton_trees), random_sample, start_idx, sample_idxes, self.is_deterministic)
        self._curr_motion_ids = sample_idxes.to(self._device)

        quats, transs, aas, nfs, fpss, bodies, heading_u, files = [], [], [], [], [], [], [], []
        randomise = not (self.is_deterministic or self.im_eval)
        J = _cabi.NUM_BODIES
        for f, idx in enumerate(sample_idxes.tolist()):
            clip = self._motion_data_list[idx]
            seq_len = clip["root_trans_offset"].shape[0]
            if self.max_length == -1 or seq_len < self.max_length:  # :773-778
                start, end = 0, seq_len
            else:
                start = 0 if self.is_deterministic else random.randint(0, seq_len - self.max_length)
                end = start + self.max_length
            nf = end - start
            if randomise:
                heading_u.append(np.random.random())  # :790
                # Reference quirk kept (:783, :794): the heading is written IN PLACE into the window of the clip's
                # own pose_aa (to_torch of a numpy slice shares memory), so it persists in the caller's data.
                aa = clip["pose_aa"]
                window = (aa.numpy() if isinstance(aa, torch.Tensor) else aa)[start:end]
                z, w = heading_half_angle(heading_u[-1])
                window[:, :3] = heading_on_rotvec(window[:, :3], z, w)
            transs.append(_as_np64(clip["root_trans_offset"])[start:end])
            quats.append(_as_np64(clip["pose_quat_global"])[start:end])
            if "beta" in clip:  # :376-381 — the FILE's pose_aa, uncropped: with a crop _motion_aa has more rows than gts
                aas.append(clip)  # read after the loop: a clip sampled twice carries both headings, as in the reference
                bodies.append(torch.as_tensor(gender_betas[f], dtype=torch.float32))
            else:
                aas.append(nf)
                bodies.append(torch.zeros(17))
            nfs.append(nf)
            fpss.append(clip.get("fps", 30))
            files.append(clip)

        built = build_motion_tensors(
            np.concatenate(quats) if quats else np.zeros((0, J, 4)), np.concatenate(transs) if transs else np.zeros((0, 3)),
            None, nfs, fpss, np.asarray(skeleton_trees[0].parent_indices),
            np.stack([np.asarray(t.local_translation, dtype=np.float32) for t in skeleton_trees]),
            heading_u=np.asarray(heading_u) if randomise else None, device=self._device,
        )  # fmt: skip
        aas = [np.zeros((a, J * 3)) if isinstance(a, int) else _as_np64(a["pose_aa"]).reshape(-1, J * 3) for a in aas]
        built["motion_aa"] = torch.from_numpy(np.concatenate(aas).astype(np.float32)).to(self._device)  # :391
        fps64 = np.asarray(fpss, dtype=np.float64)
        built.update(
            motion_lengths=torch.tensor((1.0 / fps64 * (np.asarray(nfs) - 1)).tolist(), dtype=torch.float32),  # :372
            motion_dt=torch.tensor((1.0 / fps64).tolist(), dtype=torch.float32),
            motion_fps=torch.tensor(fps64.tolist(), dtype=torch.float32),
            motion_bodies=torch.stack(bodies),
            motion_limb_weights=(limb_weights.detach().clone() if isinstance(limb_weights, torch.Tensor)
                                 else torch.tensor(np.array(limb_weights))).to(torch.float32),
        )
        if self._handle:
            self.__del__()
        MotionLib.__init__(self, built, device=self._device)
        self.grvs, self.gravs = self.gvs[:, 0], self.gavs[:, 0]  # global_root_velocity / _angular_velocity (:399-400)
        self.num_joints = len(skeleton_trees[0].node_names)
        return files
