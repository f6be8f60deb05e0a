"""Seeded synthetic inputs for the PHC step path (no Isaac Gym / AMASS / SMPL files).

The generators emit exactly the tensor set the reference's loader leaves on the
device (``MotionLibBase.load_motions``, PHC/motion_lib.py:396-420) and the sim
state the env wraps from PhysX (AoS, 13 floats per rigid body,
PHC/envs/humanoid_phc.py:542-549).  They are data plumbing, written with torch ops
so the same code fills a CPU fixture or 180 GB of HBM; nothing here is on the
measured path.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

NUM_BODIES = 24  # SMPL humanoid, PHC/body_sets.py:11-36
NUM_DOF_JOINTS = 23
SIM_DT = 2 * (1.0 / 60.0)  # control_freq_inv * sim_timestep, PHC/envs/isaacgym_env.py:39-41


@dataclass
class MotionData:
    """The A0 tensor set of SURVEY §8(a): frames of all clips concatenated on dim 0."""

    gts: torch.Tensor  # [F,24,3] global translation
    grs: torch.Tensor  # [F,24,4] global rotation, xyzw
    lrs: torch.Tensor  # [F,24,4] local rotation, xyzw
    gvs: torch.Tensor  # [F,24,3] global linear velocity
    gavs: torch.Tensor  # [F,24,3] global angular velocity
    dvs: torch.Tensor  # [F,23,3] dof velocity
    motion_aa: torch.Tensor  # [F,72]
    motion_lengths: torch.Tensor  # [M] f32 seconds
    motion_num_frames: torch.Tensor  # [M] i64
    motion_dt: torch.Tensor  # [M] f32
    motion_fps: torch.Tensor  # [M] f32
    length_starts: torch.Tensor  # [M] i64 exclusive prefix sum of num_frames
    motion_bodies: torch.Tensor  # [M,17]
    motion_limb_weights: torch.Tensor  # [M,10]

    def to(self, device) -> "MotionData":
        return MotionData(**{k: v.to(device) for k, v in self.__dict__.items()})

    def as_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.__dict__)

    @property
    def num_motions(self) -> int:
        return int(self.motion_num_frames.shape[0])

    @property
    def total_frames(self) -> int:
        return int(self.gts.shape[0])


def _rotvec_to_quat(r: torch.Tensor) -> torch.Tensor:
    ang = r.norm(dim=-1, keepdim=True)
    half = 0.5 * ang
    k = torch.where(ang > 1e-8, torch.sin(half) / ang.clamp_min(1e-8), torch.full_like(ang, 0.5))
    q = torch.cat([r * k, torch.cos(half)], dim=-1)
    return q / q.norm(dim=-1, keepdim=True)


def _smooth_rotations(
    frame_phase: torch.Tensor,  # [F] frame index inside its clip, float
    clip_of_frame: torch.Tensor,  # [F] i64
    num_motions: int,
    gen: torch.Generator,
    step_lo: float,
    step_hi: float,
    base_scale: float,
    flip_frac: float,
) -> torch.Tensor:
    """Per-(clip, body) rotation-vector trajectory r(f) = r0 + w*f + a*sin(k f + p).

    Consecutive frames differ by a rotation of about |w| rad, i.e. an inter-frame
    slerp half-angle of |w|/2; |w|/2 is drawn log-uniformly from [step_lo, step_hi]
    (the mocap-like regime of SURVEY §8(d)).  ``flip_frac`` of the quaternions get
    their sign flipped so slerp's dot<0 branch is exercised.
    """
    dev = frame_phase.device
    M, J = num_motions, NUM_BODIES

    def rnd(*shape):
        return torch.rand(*shape, generator=gen, device=dev)

    def rndn(*shape):
        return torch.randn(*shape, generator=gen, device=dev)

    r0 = rndn(M, J, 3) * base_scale
    w_dir = rndn(M, J, 3)
    w_dir = w_dir / w_dir.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    half_step = torch.exp(rnd(M, J, 1) * (math.log(step_hi) - math.log(step_lo)) + math.log(step_lo))
    w = w_dir * (2.0 * half_step)
    amp = rndn(M, J, 3) * 0.2
    k = (rnd(M, J, 1) * 0.15 + 0.02)
    ph = rnd(M, J, 1) * (2 * math.pi)

    f = frame_phase.view(-1, 1, 1)
    c = clip_of_frame
    r = r0[c] + w[c] * f + amp[c] * torch.sin(k[c] * f + ph[c])
    q = _rotvec_to_quat(r)
    if flip_frac > 0:
        flip = rnd(q.shape[0], J, 1) < flip_frac
        q = torch.where(flip, -q, q)
    return q.contiguous()


def make_motion_lib(
    num_motions: int,
    min_frames: int = 60,
    max_frames: int = 300,
    fps_choices=(30,),
    seed: int = 1234,
    device="cpu",
    rot_regime: str = "mocap",
    flip_frac: float = 0.05,
    frames_per_motion: Optional[torch.Tensor] = None,
) -> MotionData:
    """Synthetic clip store with the reference's layout and dtypes.

    ``rot_regime`` = "mocap" (smooth, inter-frame half-angle 0.005-0.1 rad) or
    "random" (i.i.d. unit quaternions per frame).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    M, J = num_motions, NUM_BODIES

    if frames_per_motion is None:
        nf = torch.randint(min_frames, max_frames + 1, (M,), generator=gen, device=dev, dtype=torch.int64)
    else:
        nf = frames_per_motion.to(dev, torch.int64)
    fps_tab = torch.tensor(list(fps_choices), dtype=torch.float64, device=dev)
    fps = fps_tab[torch.randint(0, len(fps_choices), (M,), generator=gen, device=dev)]
    # curr_dt = 1.0 / fps ; curr_len = 1.0 / fps * (num_frames - 1) in float64, then cast
    # (PHC/motion_lib.py:375-396)
    dt64 = 1.0 / fps
    len64 = dt64 * (nf - 1).to(torch.float64)
    starts = torch.cumsum(nf, 0) - nf
    F = int(nf.sum().item())

    clip_of_frame = torch.repeat_interleave(torch.arange(M, device=dev), nf)
    frame_in_clip = (torch.arange(F, device=dev) - starts[clip_of_frame]).to(torch.float32)

    def rndn(*shape):
        return torch.randn(*shape, generator=gen, device=dev)

    if rot_regime == "mocap":
        grs = _smooth_rotations(frame_in_clip, clip_of_frame, M, gen, 0.005, 0.1, 1.5, flip_frac)
        lrs = _smooth_rotations(frame_in_clip, clip_of_frame, M, gen, 0.002, 0.05, 0.4, flip_frac)
    elif rot_regime == "random":
        grs = rndn(F, J, 4)
        grs = grs / grs.norm(dim=-1, keepdim=True)
        lrs = rndn(F, J, 4)
        lrs = lrs / lrs.norm(dim=-1, keepdim=True)
    else:
        raise ValueError(rot_regime)

    # root: smooth planar drift + bob around 0.9 m; bodies: fixed offsets + small sway
    root_v = rndn(M, 1, 3) * torch.tensor([0.8, 0.8, 0.0], device=dev)
    root0 = rndn(M, 1, 3) * torch.tensor([0.5, 0.5, 0.0], device=dev) + torch.tensor([0.0, 0.0, 0.9], device=dev)
    body_off = rndn(M, J, 3) * 0.3
    body_off[:, 0] = 0
    sway = rndn(M, J, 3) * 0.05
    sk = torch.rand(M, J, 1, generator=gen, device=dev) * 0.2 + 0.05
    t = frame_in_clip.view(-1, 1, 1) * dt64[clip_of_frame].to(torch.float32).view(-1, 1, 1)
    f = frame_in_clip.view(-1, 1, 1)
    c = clip_of_frame
    gts = root0[c] + root_v[c] * t + body_off[c] + sway[c] * torch.sin(sk[c] * f)

    gvs = rndn(F, J, 3)
    gavs = rndn(F, J, 3)
    dvs = rndn(F, NUM_DOF_JOINTS, 3)
    motion_aa = rndn(F, J * 3) * 0.5

    return MotionData(
        gts=gts.float().contiguous(),
        grs=grs.float().contiguous(),
        lrs=lrs.float().contiguous(),
        gvs=gvs,
        gavs=gavs,
        dvs=dvs,
        motion_aa=motion_aa,
        motion_lengths=len64.to(torch.float32),
        motion_num_frames=nf,
        motion_dt=dt64.to(torch.float32),
        motion_fps=fps.to(torch.float32),
        length_starts=starts.to(torch.int64),
        motion_bodies=rndn(M, 17),
        motion_limb_weights=torch.rand(M, 10, generator=gen, device=dev),
    )


@dataclass
class Clock:
    """Per-env motion clock held by the env."""

    progress_buf: torch.Tensor  # [N] int16
    motion_start_times: torch.Tensor  # [N] f32
    motion_start_times_offset: torch.Tensor  # [N] f32
    global_offset: torch.Tensor  # [N,3] f32
    sampled_motion_ids: torch.Tensor  # [N] i64

    def to(self, device) -> "Clock":
        return Clock(**{k: v.to(device) for k, v in self.__dict__.items()})

    def clone(self) -> "Clock":
        return Clock(**{k: v.clone() for k, v in self.__dict__.items()})


def make_clock(
    lib: MotionData,
    num_envs: int,
    seed: int = 4321,
    ids: str = "mod",
    aligned: bool = True,
    max_progress: int = 0,
    device=None,
) -> Clock:
    """Start times in the reference's ``sample_time_interval`` regime
    (PHC/motion_lib.py:526-535): multiples of 1/30 s with k ~ U{0..nf-2}; or, with
    ``aligned=False``, uniform in [0, len).  ``ids``: "mod" = arange % M (== arange
    when M == N, the reference regime), "random" = U{0..M-1} (config 4)."""
    dev = torch.device(device) if device is not None else lib.gts.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    M = lib.num_motions
    if ids == "mod":
        mids = torch.arange(num_envs, device=dev, dtype=torch.int64) % M
    elif ids == "random":
        mids = torch.randint(0, M, (num_envs,), generator=gen, device=dev, dtype=torch.int64)
    else:
        raise ValueError(ids)
    nf = lib.motion_num_frames.to(dev)[mids]
    mlen = lib.motion_lengths.to(dev)[mids]
    u = torch.rand(num_envs, generator=gen, device=dev)
    if aligned:
        k = torch.minimum((u * (nf - 1).float()).long(), (nf - 2).clamp_min(0))
        curr = 1 / 30
        start = (k * curr).to(torch.float32)  # ((phase*len)/curr).long() * curr, motion_lib.py:532-533
    else:
        start = (u * mlen).to(torch.float32)
    off = torch.randn(num_envs, 3, generator=gen, device=dev)
    off[:, 2] = 0
    if max_progress > 0:
        prog = torch.randint(0, max_progress + 1, (num_envs,), generator=gen, device=dev).to(torch.int16)
    else:
        prog = torch.zeros(num_envs, dtype=torch.int16, device=dev)
    return Clock(
        progress_buf=prog,
        motion_start_times=start,
        motion_start_times_offset=torch.zeros(num_envs, dtype=torch.float32, device=dev),
        global_offset=off.float(),
        sampled_motion_ids=mids,
    )


def _small_rotation(n: int, j: int, sigma: float, gen: torch.Generator, dev) -> torch.Tensor:
    return _rotvec_to_quat(torch.randn(n, j, 3, generator=gen, device=dev) * sigma)


def _qmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        dim=-1,
    )


def make_sim_state(
    ref: Dict[str, torch.Tensor],
    seed: int = 777,
    pos_sigma_lo: float = 0.005,
    pos_sigma_hi: float = 0.2,
    rot_sigma: float = 0.1,
    vel_sigma: float = 0.5,
    bodies_per_env: int = NUM_BODIES,
) -> torch.Tensor:
    """AoS sim state ``[N, bodies_per_env, 13]`` = reference state + noise.

    ``ref`` holds ``rg_pos [N,24,3]``, ``rb_rot [N,24,4]``, ``body_vel``, ``body_ang_vel``
    at the reward time.  Per-env position sigma is log-uniform in
    [pos_sigma_lo, pos_sigma_hi] so termination flags are mixed and some envs sit near
    the 0.25 m threshold; rotations are ref (x) small random rotation, renormalised.
    """
    pos, rot, vel, ang = ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"]
    dev = pos.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    N, J = pos.shape[0], pos.shape[1]
    u = torch.rand(N, 1, 1, generator=gen, device=dev)
    sig = torch.exp(u * (math.log(pos_sigma_hi) - math.log(pos_sigma_lo)) + math.log(pos_sigma_lo))
    state = torch.zeros(N, bodies_per_env, 13, dtype=torch.float32, device=dev)
    state[:, :J, 0:3] = pos + torch.randn(N, J, 3, generator=gen, device=dev) * sig
    q = _qmul(rot, _small_rotation(N, J, rot_sigma, gen, dev))
    state[:, :J, 3:7] = q / q.norm(dim=-1, keepdim=True)
    state[:, :J, 7:10] = vel + torch.randn(N, J, 3, generator=gen, device=dev) * vel_sigma
    state[:, :J, 10:13] = ang + torch.randn(N, J, 3, generator=gen, device=dev) * vel_sigma
    return state


def body_views(state: torch.Tensor, num_bodies: int = NUM_BODIES):
    """The four stride-13 views the env hands to the path (humanoid_phc.py:546-549)."""
    return (
        state[..., :num_bodies, 0:3],
        state[..., :num_bodies, 3:7],
        state[..., :num_bodies, 7:10],
        state[..., :num_bodies, 10:13],
    )


def reward_time(clock: Clock, dt: float = SIM_DT, extra_steps: int = 0) -> torch.Tensor:
    """t = progress*dt + start + start_offset (humanoid_phc.py:1236), fp32 op order kept."""
    return (clock.progress_buf + extra_steps) * dt + clock.motion_start_times + clock.motion_start_times_offset


def make_case(
    num_envs: int,
    num_motions: int,
    query: Callable[[MotionData, torch.Tensor, torch.Tensor, torch.Tensor], Dict[str, torch.Tensor]],
    seed: int = 1234,
    device="cpu",
    min_frames: int = 60,
    max_frames: int = 300,
    fps_choices=(30,),
    ids: str = "mod",
    aligned: bool = True,
    rot_regime: str = "mocap",
    sim_at_progress: int = 1,
    max_progress: int = 0,
):
    """A whole synthetic workload: library, clock and a sim state that matches the
    reference pose at ``progress + sim_at_progress`` (the reward time of the next step).
    ``query(lib, ids, times, offset)`` supplies the reference state — the oracle in
    tests, the CUDA kernel in the bench."""
    lib = make_motion_lib(
        num_motions, min_frames, max_frames, fps_choices, seed=seed, device=device, rot_regime=rot_regime
    )
    clock = make_clock(lib, num_envs, seed=seed + 1, ids=ids, aligned=aligned, max_progress=max_progress)
    t = reward_time(clock, extra_steps=sim_at_progress)
    ref = query(lib, clock.sampled_motion_ids, t, clock.global_offset)
    state = make_sim_state(ref, seed=seed + 2)
    return lib, clock, state
