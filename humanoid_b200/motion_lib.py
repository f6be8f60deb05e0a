"""HBM-resident motion library with the reference's attribute and method names.

Drop-in for the query side of ``MotionLibBase`` (PHC/motion_lib.py:549-673): the object
holds the same tensors under the same names (``gts, grs, lrs, gvs, gavs, dvs, _motion_aa,
_motion_lengths, _motion_num_frames, _motion_dt, length_starts, _motion_bodies,
_motion_limb_weights``) and ``get_motion_state`` returns the same 13-key dict, computed by
one CUDA kernel (``phc_motion_state``) instead of ~164 ATen launches.  Loading clips from
AMASS pkl files (motion_lib.py:180-430) is out of scope; the tensors are handed in.
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _cabi


class MotionLib:
    def __init__(self, data, device=None):
        """``data``: a ``synth.MotionData`` or a dict with the A0 tensor set."""
        d = data.as_dict() if hasattr(data, "as_dict") else dict(data)
        dev = torch.device(device) if device is not None else d["gts"].device
        if dev.type != "cuda":
            raise _cabi.PhcError("MotionLib lives in HBM: pass device='cuda' (there is no CPU path)")
        self._device = dev

        def put(key, dtype):
            return d[key].to(device=dev, dtype=dtype).contiguous()

        self.gts = put("gts", torch.float32)
        self.grs = put("grs", torch.float32)
        self.lrs = put("lrs", torch.float32)
        self.gvs = put("gvs", torch.float32)
        self.gavs = put("gavs", torch.float32)
        self.dvs = put("dvs", torch.float32)
        self._motion_aa = put("motion_aa", torch.float32)
        self._motion_lengths = put("motion_lengths", torch.float32)
        self._motion_num_frames = put("motion_num_frames", torch.int64)
        self._motion_dt = put("motion_dt", torch.float32)
        self.length_starts = put("length_starts", torch.int64)
        self._motion_bodies = put("motion_bodies", torch.float32)
        self._motion_limb_weights = put("motion_limb_weights", torch.float32)
        if "motion_fps" in d:
            self._motion_fps = put("motion_fps", torch.float32)
        self._num_motions = int(self._motion_num_frames.shape[0])
        self.num_bodies = int(self.gts.shape[1])
        self.motion_ids = torch.arange(self._num_motions, dtype=torch.long, device=dev)
        if self.num_bodies != _cabi.NUM_BODIES:
            raise _cabi.PhcError(f"expected {_cabi.NUM_BODIES} bodies, got {self.num_bodies}")

        lib = _cabi.load()
        desc = _cabi.PhcLibDesc(
            self.gts.data_ptr(), self.grs.data_ptr(), self.lrs.data_ptr(), self.gvs.data_ptr(),
            self.gavs.data_ptr(), self.dvs.data_ptr(), self._motion_aa.data_ptr(),
            self._motion_lengths.data_ptr(), self._motion_num_frames.data_ptr(), self._motion_dt.data_ptr(),
            self.length_starts.data_ptr(), self._motion_bodies.data_ptr(), self._motion_limb_weights.data_ptr(),
            int(self.gts.shape[0]), self._num_motions,
        )  # fmt: skip
        handle = C.c_void_p()
        _cabi.check(lib.phc_lib_create(C.byref(desc), C.byref(handle)), "phc_lib_create")
        self._handle = handle
        self.repack()

    def repack(self):
        """Build (or rebuild, after the frame tensors were rewritten in place) the packed
        [F, 312] frame table the fused step's TMA path reads."""
        _cabi.check(_cabi.load().phc_lib_pack(self._handle, _cabi.stream_ptr(self._device)), "phc_lib_pack")

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _cabi.load().phc_lib_destroy(h)
            except Exception:
                pass
            self._handle = None

    # -- reference accessors (motion_lib.py:432-452, 537-547) ---------------------------
    def num_motions(self) -> int:
        return self._num_motions

    def get_total_length(self):
        return float(self._motion_lengths.sum())

    def get_motion_length(self, motion_ids=None):
        return self._motion_lengths if motion_ids is None else self._motion_lengths[motion_ids]

    def _get_num_bodies(self) -> int:
        return self.num_bodies

    def get_motion_num_steps(self, motion_ids=None):  # motion_lib.py:543-547
        sim_fps = getattr(self, "_sim_fps", 30.0)
        nf = self._motion_num_frames if motion_ids is None else self._motion_num_frames[motion_ids]
        fps = self._motion_fps if motion_ids is None else self._motion_fps[motion_ids]
        return (nf * sim_fps / fps).ceil().int()

    def sample_time(self, motion_ids, truncate_time=None, phase=None):  # motion_lib.py:515-524
        if phase is None:
            phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len = motion_len - truncate_time
        return phase * motion_len

    def sample_time_interval(self, motion_ids, truncate_time=None, phase=None):  # motion_lib.py:526-535
        """Start times on the 30 Hz grid.  The env's reset draws these inside ``phc_reset_envs``; this is the
        standalone method, a handful of elementwise torch ops on the device."""
        if phase is None:
            phase = torch.rand(motion_ids.shape, device=self._device)
        motion_len = self._motion_lengths[motion_ids]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len = motion_len - truncate_time
        curr_fps = 1 / 30
        return ((phase * motion_len) / curr_fps).long() * curr_fps

    @property
    def handle(self) -> C.c_void_p:
        return self._handle

    # -- get_root_pos_smpl (motion_lib.py:628-653) -------------------------------------------
    def get_root_pos_smpl(self, motion_ids, motion_times):
        """Blended root position without the global offset: the same ``(1-b)*p0 + b*p1`` on body 0 that
        ``get_motion_state`` returns as ``root_pos`` (called once per motion resample, humanoid_phc.py:1377)."""
        return {"root_pos": self.get_motion_state(motion_ids, motion_times)["root_pos"]}

    # -- _calc_frame_blend (motion_lib.py:655-665) -----------------------------------------
    def _calc_frame_blend(self, time, len, num_frames, dt):  # noqa: A002 - the reference's names
        for t, nm, dty in ((time, "time", torch.float32), (len, "len", torch.float32),
                           (num_frames, "num_frames", torch.int64), (dt, "dt", torch.float32)):  # fmt: skip
            _cabi.require_cuda(t, nm, dty)
        n = time.numel()
        time, len, num_frames, dt = (x.contiguous() for x in (time, len, num_frames, dt))
        idx0 = torch.empty(n, dtype=torch.int64, device=time.device)
        idx1 = torch.empty_like(idx0)
        blend = torch.empty(n, dtype=torch.float32, device=time.device)
        _cabi.check(
            _cabi.load().phc_calc_frame_blend(
                time.data_ptr(), len.data_ptr(), num_frames.data_ptr(), dt.data_ptr(), n,
                idx0.data_ptr(), idx1.data_ptr(), blend.data_ptr(), _cabi.stream_ptr(time.device),
            ),
            "phc_calc_frame_blend",
        )  # fmt: skip
        return idx0.view(time.shape), idx1.view(time.shape), blend.view(time.shape)

    # -- get_motion_state (motion_lib.py:549-626) ---------------------------------------------
    def get_motion_state(
        self, motion_ids: torch.Tensor, motion_times: torch.Tensor, offset: Optional[torch.Tensor] = None,
        with_frame_info: bool = False,
    ) -> Dict[str, torch.Tensor]:  # fmt: skip
        _cabi.require_cuda(motion_ids, "motion_ids", torch.int64)
        _cabi.require_cuda(motion_times, "motion_times")
        if motion_times.dtype == torch.float64:
            raise _cabi.PhcError("motion_times must not be float64 (motion_lib.py:589-590)")
        motion_ids = motion_ids.contiguous()
        motion_times = motion_times.to(torch.float32).contiguous()
        n = motion_ids.numel()
        if motion_times.numel() != n:
            raise _cabi.PhcError("motion_ids and motion_times differ in length")
        if offset is not None:
            _cabi.require_cuda(offset, "offset", torch.float32)
            if offset.shape != (n, 3):
                raise _cabi.PhcError(f"offset must be [{n},3], got {tuple(offset.shape)}")
            offset = offset.contiguous()
        dev = motion_ids.device
        J = self.num_bodies

        def f32(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        res = {
            "root_pos": f32(n, 3),
            "root_rot": f32(n, 4),
            "dof_pos": f32(n, (J - 1) * 3),
            "root_vel": f32(n, 3),
            "root_ang_vel": f32(n, 3),
            "dof_vel": f32(n, (J - 1) * 3),
            "motion_aa": f32(n, J * 3),
            "rg_pos": f32(n, J, 3),
            "rb_rot": f32(n, J, 4),
            "body_vel": f32(n, J, 3),
            "body_ang_vel": f32(n, J, 3),
            "motion_bodies": f32(n, self._motion_bodies.shape[1]),
            "motion_limb_weights": f32(n, self._motion_limb_weights.shape[1]),
        }
        extra = {}
        if with_frame_info:
            extra = {
                "frame_idx0": torch.empty(n, dtype=torch.int64, device=dev),
                "frame_idx1": torch.empty(n, dtype=torch.int64, device=dev),
                "blend": f32(n),
            }
        out = _cabi.PhcMotionOut(
            *[res[k].data_ptr() for k in (
                "root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "motion_aa", "rg_pos",
                "rb_rot", "body_vel", "body_ang_vel", "motion_bodies", "motion_limb_weights")],
            *[(extra[k].data_ptr() if with_frame_info else None) for k in ("frame_idx0", "frame_idx1", "blend")],
        )  # fmt: skip
        _cabi.check(
            _cabi.load().phc_motion_state(
                self._handle, motion_ids.data_ptr(), motion_times.data_ptr(), _cabi.ptr(offset), n,
                C.byref(out), _cabi.stream_ptr(dev),
            ),
            "phc_motion_state",
        )  # fmt: skip
        res.update(extra)
        return res
