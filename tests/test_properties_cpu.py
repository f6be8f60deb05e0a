"""CPU property tests (hypothesis) of the oracle's integer path and slerp — against the live
reference where /root/reference exists, and against invariants everywhere."""

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import phc_oracle as O
from oracle import ref_loader

f32 = np.float32
SIM_DT = 2 * (1.0 / 60.0)


@st.composite
def clip_and_time(draw):
    fps = draw(st.sampled_from([30.0, 60.0, 120.0]))
    nf = draw(st.integers(min_value=2, max_value=7000))
    dt = 1.0 / fps
    length = dt * (nf - 1)
    kind = draw(st.sampled_from(["boundary", "env", "uniform", "outside"]))
    if kind == "boundary":  # exact frame boundaries and their float neighbours
        k = draw(st.integers(min_value=0, max_value=nf))
        t = f32(k * dt)
        t = np.nextafter(t, f32(draw(st.sampled_from([-np.inf, 0.0, np.inf])))) if draw(st.booleans()) else t
    elif kind == "env":  # progress*dt + k/30, the env's own arithmetic (humanoid_phc.py:1236)
        p = draw(st.integers(min_value=0, max_value=400))
        k = draw(st.integers(min_value=0, max_value=nf))
        t = float((torch.tensor([p], dtype=torch.int16) * SIM_DT + torch.tensor([k * (1 / 30)], dtype=torch.float32))[0])
    elif kind == "uniform":
        t = f32(draw(st.floats(min_value=0.0, max_value=float(length), allow_subnormal=False)))
    else:
        t = f32(draw(st.floats(min_value=-2.0, max_value=float(length) * 3 + 1, allow_subnormal=False)))
    return float(t), float(length), nf, dt


@settings(max_examples=400, deadline=None)
@given(clip_and_time())
def test_frame_blend_invariants(x):
    t, length, nf, dt = x
    i0, i1, bl = O.frame_blend(torch.tensor([t], dtype=torch.float32), torch.tensor([length], dtype=torch.float32),
                               torch.tensor([nf]), torch.tensor([dt], dtype=torch.float32))  # fmt: skip
    assert 0 <= int(i0) <= nf - 1 and int(i1) == min(int(i0) + 1, nf - 1)
    assert 0.0 <= float(bl) <= 1.0
    if t <= 0:
        assert int(i0) == 0 and float(bl) == 0.0
    if t >= length * 1.001:
        assert int(i0) == nf - 1


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@settings(max_examples=300, deadline=None)
@given(st.lists(clip_and_time(), min_size=1, max_size=64))
def test_frame_blend_bit_exact_vs_reference(xs):
    _, _, ref_ml = ref_loader.load()
    t = torch.tensor([x[0] for x in xs], dtype=torch.float32)
    ln = torch.tensor([x[1] for x in xs], dtype=torch.float32)
    nf = torch.tensor([x[2] for x in xs])
    dt = torch.tensor([x[3] for x in xs], dtype=torch.float32)
    want = object.__new__(ref_ml.MotionLibBase)._calc_frame_blend(t, ln, nf, dt)
    got = O.frame_blend(t, ln, nf, dt)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


unit = st.floats(min_value=-1.0, max_value=1.0, allow_subnormal=False)


@settings(max_examples=300, deadline=None)
@given(st.tuples(unit, unit, unit, unit), st.tuples(unit, unit, unit, unit),
       st.floats(min_value=0.0, max_value=1.0, allow_subnormal=False), st.floats(min_value=1e-6, max_value=1.0))  # fmt: skip
def test_slerp_properties(a, b, t, eps):
    q0 = torch.tensor(a, dtype=torch.float32)
    q1 = torch.tensor(b, dtype=torch.float32)
    if float(q0.norm()) < 1e-3 or float(q1.norm()) < 1e-3:
        return
    q0, q1 = q0 / q0.norm(), q1 / q1.norm()
    tt = torch.tensor([t], dtype=torch.float32)
    out = O.slerp(q0[None], q1[None], tt[None])[0]
    assert not torch.isnan(out).any()
    # endpoints, sign symmetry of q1, and near-unit norm away from the averaging branch
    assert torch.allclose(O.slerp(q0[None], q1[None], torch.zeros(1, 1))[0], q0, atol=1e-3)
    if abs(float((q0 * q1).sum())) > 1e-3:  # the flip is decided by the sign of the dot product
        assert torch.allclose(O.slerp(q0[None], (-q1)[None], tt[None])[0], out, atol=1e-6)
    assert 0.99 < float(out.norm()) < 1.01
    # identical inputs hit the |cos| >= 1 / tiny-sin guards and must return q0
    same = O.slerp(q0[None], q0[None], tt[None])[0]
    assert torch.allclose(same, q0, atol=1e-6)
    if ref_loader.available():
        ref_tu, _, _ = ref_loader.load()
        assert torch.equal(ref_tu.slerp(q0[None], q1[None], tt[None])[0], out)
