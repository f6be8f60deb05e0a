"""CPU: the C-ABI library loads, exports every symbol include/phc_b200.h declares, and
rejects bad arguments with error codes before touching a device.  No compute calls."""

import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from humanoid_b200 import _cabi

HEADER = os.path.join(ROOT, "include", "phc_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _cabi.load()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"PHC_API[^;(]*?\b(phc_\w+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == names


def test_abi_version_and_strerror(lib):
    assert lib.phc_abi_version() == _cabi.ABI_VERSION
    assert lib.phc_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert lib.phc_strerror(code) not in (b"ok", b"unknown error")
    assert lib.phc_strerror(-99) == b"unknown error"


def test_struct_layouts_match_the_header():
    # sizes follow from the header's field lists on LP64
    assert C.sizeof(_cabi.PhcLibDesc) == 13 * 8 + 2 * 8
    assert C.sizeof(_cabi.PhcView) == 24
    assert C.sizeof(_cabi.PhcBodyState) == 4 * 24 + 8
    assert C.sizeof(_cabi.PhcMotionOut) == 16 * 8
    assert C.sizeof(_cabi.PhcRewardSpec) == 32
    assert C.sizeof(_cabi.PhcHostStepArgs) == 11 * 8
    fields = [f[0] for f in _cabi.PhcStepArgs._fields_]
    src = open(HEADER).read()
    body = src[src.index("typedef struct PhcStepArgs {") : src.index("} PhcStepArgs;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)  # drop comments, then take each declarator
    in_header = re.findall(r"\b(\w+)\s*;", body)
    assert fields == in_header, (fields, in_header)


def test_every_struct_matches_the_header_as_a_c_compiler_lays_it_out(tmp_path):
    """gcc compiles a C99 consumer of include/phc_b200.h (so the header has no C++ / torch types in it) that
    prints sizeof and offsetof for every field the ctypes mirror declares; both sides must agree, field by
    field — the check that catches an argument struct growing on one side only."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = [v for v in vars(_cabi).values() if isinstance(v, type) and issubclass(v, C.Structure) and v is not C.Structure]
    assert len(structs) >= 12
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "phc_b200.h"', "int main(void) {"]
    for st in structs:
        lines.append(f'  printf("{st.__name__} %zu\\n", sizeof({st.__name__}));')
        for name, *_ in st._fields_:
            lines.append(f'  printf("{st.__name__}.{name} %zu\\n", offsetof({st.__name__}, {name}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.dirname(HEADER),
                    str(src), "-o", str(exe)], check=True)  # fmt: skip
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for st in structs:
        assert int(out[st.__name__]) == C.sizeof(st), st.__name__
        for name, *_ in st._fields_:
            assert int(out[f"{st.__name__}.{name}"]) == getattr(st, name).offset, f"{st.__name__}.{name}"


def test_argument_validation_returns_codes(lib):
    h = C.c_void_p()
    assert lib.phc_lib_create(None, C.byref(h)) == -1
    d = _cabi.PhcLibDesc()
    assert lib.phc_lib_create(C.byref(d), C.byref(h)) == -1  # NULL tensors
    for k in ("gts", "grs", "lrs", "gvs", "gavs", "motion_lengths", "motion_num_frames", "motion_dt", "length_starts"):
        setattr(d, k, 4096)
    d.total_frames, d.num_motions = 0, 1
    assert lib.phc_lib_create(C.byref(d), C.byref(h)) == -2  # empty library
    d.total_frames = 10
    d.gts = 4100  # not 16-B aligned
    assert lib.phc_lib_create(C.byref(d), C.byref(h)) == -3
    d.gts = 4096
    assert lib.phc_lib_create(C.byref(d), C.byref(h)) == 0 and h.value
    # zero-sized batches are no-ops; negative sizes and NULL buffers are errors (no launch happens)
    assert lib.phc_motion_state(h, None, None, None, 0, None, None) == 0
    assert lib.phc_motion_state(h, None, None, None, -1, None, None) == -2
    assert lib.phc_motion_state(h, None, None, None, 4, None, None) == -1
    assert lib.phc_step_fused(h, None, 8, None) == -1
    a = _cabi.PhcStepArgs()
    assert lib.phc_step_fused(h, C.byref(a), 8, None) == -1
    assert lib.phc_calc_frame_blend(None, None, None, None, 5, None, None, None, None) == -1
    assert lib.phc_obs_moments(None, 4, 4, 2, None, None) == -2  # stride < cols
    assert lib.phc_imitation_obs(None, 0, None, 0, None, None, 3, 0, 1, 6, None, 0, None) == -2  # T < 1
    # phc_reset_envs: the StateInit modes are validated before anything is launched
    r = _cabi.PhcResetArgs()
    V = _cabi.PhcView
    r.body = _cabi.PhcBodyState(V(4096, 312, 13), V(4096 + 12, 312, 13), V(4096 + 28, 312, 13), V(4096 + 40, 312, 13), 24)
    for k in ("progress_buf", "reset_buf", "terminate_buf", "motion_start_times", "motion_start_times_offset",
              "sampled_motion_ids", "obs_buf", "humanoid_root_states", "dof_pos", "dof_vel", "env_mask"):
        setattr(r, k, 4096)
    r.root_stride, r.dof_stride, r.dof_elem_stride, r.obs_stride, r.time_steps, r.dt = 13, 138, 2, 934, 1, 1 / 30
    r.state_init = 7
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -4  # no such StateInit
    r.state_init = _cabi.STATE_INIT_RANDOM
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -1  # RANDOM needs the caller's uniform numbers
    r.state_init = _cabi.STATE_INIT_HYBRID
    r.phase = 4096
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -1  # HYBRID needs the Bernoulli mask ...
    r.default_mask = 4096
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -1  # ... and the initial buffers
    r.initial_root_states = r.initial_dof_pos = r.initial_dof_vel = 4096
    r.initial_root_stride, r.initial_dof_stride = 13, 60
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -2  # 69 dofs per row
    r.state_init = _cabi.STATE_INIT_DEFAULT
    r.default_mask = None
    assert lib.phc_reset_envs(h, C.byref(r), 8, None) == -2
    assert lib.phc_reset_envs(h, C.byref(r), -1, None) == -2
    lib.phc_lib_destroy(h)
    lib.phc_lib_destroy(None)


def test_python_layer_refuses_cpu_tensors():
    from humanoid_b200 import MotionLib, compute_imitation_reward, synth

    data = synth.make_motion_lib(2, 4, 6)
    with pytest.raises(_cabi.PhcError):
        MotionLib(data)  # CPU tensors, no device given
    z = torch.zeros(4, 24, 13)
    p, r, v, a = synth.body_views(z)
    with pytest.raises(_cabi.PhcError):
        compute_imitation_reward(p[:, 0], r[:, 0], p, r, v, a, p, r, v, a, dict(
            k_pos=1, k_rot=1, k_vel=1, k_ang_vel=1, w_pos=1, w_rot=1, w_vel=1, w_ang_vel=1))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "humanoid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "phc_oracle" not in src, f"{f} mentions the oracle module"
