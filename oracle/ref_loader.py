"""Import the reference's own hot-path functions.  TEST / BASELINE INFRASTRUCTURE — never the product path.

Two places can hold them: ``/root/reference`` (the build container) and the staged, git-ignored copy of the
reference's pure-Python package under ``baseline/_ref/`` that ``__graft_entry__.build()`` makes in the container
and that travels to the GPU box with the snapshot (the offline ``pip install --target baseline/_ref`` of the
contract fails here: the package's build backend, hatchling, is not in the wheelhouse; what the wheel would
install — the ``puffer_phc`` package directory — is copied instead).  Users: ``tests/golden/make_golden.py``
(fixture generation), the tests that pin ``oracle/phc_oracle.py`` directly against the reference, and
``bench.py``'s CPU legs (``--impl reference`` and ``cpu_baseline``, ``kind: "reference"``): the reference's
UNMODIFIED functions timed on the box's host cores.  Nothing marked ``gpu`` and no product module imports this.

Recipe: put ``packages/puffer-phc`` on sys.path, stub the three
``smpl_sim`` modules that ``puffer_phc.motion_lib`` imports at top level
(motion_lib.py:51), and build a ``MotionLibBase`` with ``object.__new__`` + attribute
injection, because its ``__init__`` needs an AMASS pkl and SMPL model files.
"""

from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE_ROOT = "/root/reference/packages/puffer-phc"
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _root():
    roots = (STAGED_ROOT,) if os.environ.get("PHC_REF_STAGED_ONLY") else (SOURCE_ROOT, STAGED_ROOT)
    for r in roots:
        if os.path.isfile(os.path.join(r, "puffer_phc", "envs", "common.py")):
            return r
    return None


REFERENCE_ROOT = _root() or SOURCE_ROOT


def available() -> bool:
    return _root() is not None


def where() -> str:
    r = _root()
    return "none" if r is None else ("/root/reference" if r == SOURCE_ROOT else "baseline/_ref (staged copy)")


def stage() -> bool:
    """Copy the reference's ``puffer_phc`` package (pure Python + its XML assets, 370 KB) to ``baseline/_ref/`` —
    what its wheel would install.  Container only; returns False where /root/reference is absent."""
    import shutil

    src = os.path.join(SOURCE_ROOT, "puffer_phc")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(STAGED_ROOT, "puffer_phc")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


def load():
    """Returns (torch_utils, common, motion_lib) modules of the reference."""
    root = _root()
    if root is None:
        raise RuntimeError("reference not present (neither /root/reference nor baseline/_ref)")
    if root not in sys.path:
        sys.path.insert(0, root)
    if "smpl_sim" not in sys.modules:
        pkg = types.ModuleType("smpl_sim")
        sub = types.ModuleType("smpl_sim.smpllib")
        leaf = types.ModuleType("smpl_sim.smpllib.smpl_parser")

        class SMPL_Parser:  # never instantiated on the query path
            def __init__(self, *a, **k):
                raise RuntimeError("SMPL_Parser stub")

        leaf.SMPL_Parser = SMPL_Parser
        pkg.smpllib = sub
        sub.smpl_parser = leaf
        sys.modules["smpl_sim"] = pkg
        sys.modules["smpl_sim.smpllib"] = sub
        sys.modules["smpl_sim.smpllib.smpl_parser"] = leaf
    import puffer_phc.torch_utils as ref_tu
    import puffer_phc.envs.common as ref_common
    import puffer_phc.motion_lib as ref_ml

    return ref_tu, ref_common, ref_ml


def make_reference_lib(data):
    """A reference ``MotionLibBase`` whose A0 attributes are the synthetic tensors."""
    _, _, ref_ml = load()
    d = data.as_dict() if hasattr(data, "as_dict") else dict(data)
    lib = object.__new__(ref_ml.MotionLibBase)
    lib._device = "cpu"
    lib.gts, lib.grs, lib.lrs = d["gts"], d["grs"], d["lrs"]
    lib.gvs, lib.gavs, lib.dvs = d["gvs"], d["gavs"], d["dvs"]
    lib._motion_aa = d["motion_aa"]
    lib._motion_lengths = d["motion_lengths"]
    lib._motion_num_frames = d["motion_num_frames"]
    lib._motion_dt = d["motion_dt"]
    lib._motion_fps = d["motion_fps"]
    lib.length_starts = d["length_starts"]
    lib._motion_bodies = d["motion_bodies"]
    lib._motion_limb_weights = d["motion_limb_weights"]
    lib.num_bodies = 24
    return lib


RWD_SPECS = dict(  # asdict(RewardConfig), PHC/config.py:38-50 — what humanoid_phc.py:1307-1311 passes
    k_pos=100.0, k_rot=10.0, k_vel=0.1, k_ang_vel=0.1, w_pos=0.5, w_rot=0.3, w_vel=0.1, w_ang_vel=0.1,
    imitation_reward_dim=4, full_body_reward=True, use_power_reward=True,
)  # fmt: skip


def reference_step(lib, state, progress, start, start_off, goff, ids, term_dist, dt, time_steps=1):
    """The post-physics half of ``HumanoidPHC.step`` (PHC/envs/humanoid_phc.py:138-149) with the reference's own
    functions, orchestrated as the env does: progress += 1; ``_compute_reward`` (query at t + reward, :1230-1271);
    ``_compute_reset`` (same t: the env's one-entry cache hits, :877-899, so no second query; :1313-1335);
    ``_compute_observations`` (smpl_max self obs :963-998, query at t+dt.., v6 task obs :1050-1123, cat :949).
    ``lib`` is ``make_reference_lib(data)``; ``state`` the AoS [N, 24, 13] sim tensor.  Returns
    ``(obs, reward, reward_raw, reset, terminated)``; ``progress`` is advanced in place."""
    import torch

    _, ref_common, _ = load()
    J = 24
    pos, rot, vel, ang = state[:, :J, 0:3], state[:, :J, 3:7], state[:, :J, 7:10], state[:, :J, 10:13]
    progress += 1
    t = progress * dt + start + start_off
    ref = lib.get_motion_state(ids, t, goff)
    reward, raw = ref_common.compute_imitation_reward(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang, ref["rg_pos"], ref["rb_rot"], ref["body_vel"], ref["body_ang_vel"],
        RWD_SPECS)  # fmt: skip
    pass_time = t >= lib._motion_lengths[ids]
    n = state.shape[0]
    reset, terminated = ref_common.compute_humanoid_im_reset(
        torch.ones(n, dtype=torch.bool), progress, torch.zeros(1), torch.zeros(4, dtype=torch.long),
        pos.clone(), ref["rg_pos"].clone(), pass_time, True, term_dist, False)  # fmt: skip
    self_obs = ref_common.compute_humanoid_observations_smpl_max(pos, rot, vel, ang, None, None, True, True, True, False, False)
    refs = [lib.get_motion_state(ids, (progress + k) * dt + start + start_off, goff) for k in range(1, time_steps + 1)]

    def stack(key):
        if len(refs) == 1:
            return refs[0][key]
        return torch.stack([r[key] for r in refs], 1).reshape((-1,) + refs[0][key].shape[1:])

    task = ref_common.compute_imitation_observations_v6(
        pos[:, 0], rot[:, 0], pos, rot, vel, ang, stack("rg_pos"), stack("rb_rot"), stack("body_vel"),
        stack("body_ang_vel"), time_steps, True)  # fmt: skip
    return torch.cat([self_obs, task], dim=-1), reward, raw, reset, terminated
