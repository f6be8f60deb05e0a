"""CPU: the plain-C restatement of the integer / flag outputs (oracle/phc_oracle_int.c) against the
reference-generated fixtures — a second, independent pin for everything that must be bit-exact."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, STEP_CASES

SO = os.path.join(ROOT, "oracle", "libphc_oracle_int.so")


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(SO)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_c_frame_blend_bit_exact(clib, golden):
    g = golden("frame_blend")
    t, ln = g.inp("time").numpy(), g.inp("len").numpy()
    nf, dt = g.inp("num_frames").numpy(), g.inp("dt").numpy()
    n = t.shape[0]
    i0, i1, bl = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32)
    clib.phc_oracle_frame_blend(C.c_int64(n), p(t), p(ln), p(nf), p(dt), p(i0), p(i1), p(bl))
    assert np.array_equal(i0, g.out("frame_idx0").numpy())
    assert np.array_equal(i1, g.out("frame_idx1").numpy())
    assert np.array_equal(bl, g.out("blend").numpy())


def test_c_sample_time_bit_exact(clib, golden):
    g = golden("sample_time")
    ph = g.inp("phase").numpy()
    ln = g.inp("motion_lengths")[g.inp("ids")].numpy().copy()
    out = np.empty_like(ph)
    clib.phc_oracle_sample_time(C.c_int64(ph.shape[0]), p(ph), p(ln), p(out))
    assert np.array_equal(out, g.out("motion_time").numpy())


@pytest.mark.parametrize("case", [c for c in STEP_CASES if c != "step_eval_reset"])
def test_c_reset_flags_bit_exact(clib, golden, case):
    g = golden(case)
    state = g.inp("state")
    pos = np.ascontiguousarray(state[:, :24, 0:3].numpy())
    ref = np.ascontiguousarray(g.out("t0.rg_pos").numpy())
    n = pos.shape[0]
    prog = g.out("progress_after").numpy()
    pass_time = g.out("pass_time").numpy().astype(np.uint8)
    term = g.inp("term_dist").numpy()
    reset, terminated = np.empty(n, np.uint8), np.empty(n, np.uint8)
    clib.phc_oracle_im_reset(C.c_int64(n), C.c_int32(24), p(pos), p(ref), p(prog), p(pass_time), p(term),
                             C.c_int32(int(bool(g.inp("early")))), p(reset), p(terminated), None)  # fmt: skip
    assert np.array_equal(reset.astype(bool), g.out("reset").numpy())
    assert np.array_equal(terminated.astype(bool), g.out("terminated").numpy())


def test_c_frame_blend_matches_torch_oracle_on_random_clocks(clib):
    from oracle import phc_oracle as O

    gen = torch.Generator().manual_seed(3)
    n = 200_000
    nf = torch.randint(2, 7000, (n,), generator=gen)
    fps = torch.tensor([30.0, 60.0, 120.0], dtype=torch.float64)[torch.randint(0, 3, (n,), generator=gen)]
    dt = (1.0 / fps).float()
    ln = ((1.0 / fps) * (nf - 1)).float()
    k = (torch.rand(n, generator=gen) * nf).long()
    pr = torch.randint(0, 400, (n,), generator=gen).to(torch.int16)
    t = pr * (2 * (1.0 / 60.0)) + (k * (1 / 30)).float()
    want = O.frame_blend(t, ln, nf, dt)
    i0, i1, bl = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32)
    clib.phc_oracle_frame_blend(C.c_int64(n), p(t.numpy()), p(ln.numpy()), p(nf.numpy()), p(dt.numpy()), p(i0), p(i1), p(bl))
    assert np.array_equal(i0, want[0].numpy()) and np.array_equal(i1, want[1].numpy())
    assert np.array_equal(bl, want[2].numpy())
