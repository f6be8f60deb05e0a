"""Import the reference's own hot-path functions from /root/reference.  TEST INFRASTRUCTURE.

Works only where /root/reference exists (the build container) — the GPU box never has
it, so nothing marked ``gpu``, ``smoke()`` or ``bench.py`` may call this.  It is used
by ``tests/golden/make_golden.py`` (fixture generation) and by the container-only
tests that pin ``oracle/phc_oracle.py`` directly against the reference.

Recipe: put ``packages/puffer-phc`` on sys.path, stub the three
``smpl_sim`` modules that ``puffer_phc.motion_lib`` imports at top level
(motion_lib.py:51), and build a ``MotionLibBase`` with ``object.__new__`` + attribute
injection, because its ``__init__`` needs an AMASS pkl and SMPL model files.
"""

from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference/packages/puffer-phc"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "puffer_phc"))


def load():
    """Returns (torch_utils, common, motion_lib) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
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
