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


def test_clip_sampler_host_logic():
    """ClipSampler without the reference: ordering, filters, weights (runs on the GPU box too)."""
    import numpy as np

    from humanoid_b200.motion_build import ClipSampler

    clips = {f"c{i}": {"pose_quat_global": np.zeros((n, 24, 4))} for i, n in enumerate([5, 9, 9, 2, 7])}
    s = ClipSampler()
    s._init_clips(clips, im_eval=True)
    assert s._motion_data_keys.tolist() == ["c1", "c2", "c4", "c0", "c3"]  # longest first, stable
    s._init_clips(clips, min_length=7)
    assert s._motion_data_keys.tolist() == ["c1", "c2", "c4"]
    s._init_clips(clips)
    assert s._num_unique_motions == 5 and torch.allclose(s._sampling_prob, torch.full((5,), 0.2))
    assert s.pick_clips(7, False, 3, None, is_deterministic=False).tolist() == [3, 4, 0, 1, 2, 3, 4]
    assert s.pick_clips(3, True, 0, [4, 4, 1], is_deterministic=False).tolist() == [4, 4, 1]  # caller-supplied ids win
    assert s.curr_motion_keys.tolist() == ["c4", "c4", "c1"] and abs(float(s._sampling_batch_prob.sum()) - 1) < 1e-6
    s.update_hard_sampling_weight(["c2"])
    assert s._sampling_prob.tolist() == [0, 0, 1, 0, 0]
    assert set(s.pick_clips(16, True, 0, None, is_deterministic=False).tolist()) == {2}
    assert s.pick_clips(2, True, 0, None, is_deterministic=True).tolist() == [0, 1]  # deterministic ignores the weights
    s.update_soft_sampling_weight(["c0", "c3"])
    assert s._sampling_prob.tolist() == [0.5, 0, 0, 0.5, 0] and s._termination_history.tolist() == [1, 0, 0, 1, 0]
    assert not s.update_sampling_prob(torch.zeros(5)) and not s.update_sampling_prob(torch.ones(4))
    s.update_hard_sampling_weight([])
    assert torch.allclose(s._sampling_prob, torch.full((5,), 0.2))


def test_running_norm_checkpoint_compatibility():
    """the accumulators stay plain fp32 buffers under the reference's names, so a reference checkpoint
    loads; pickling the module drops the peer mailbox handle instead of failing."""
    import io
    import pickle

    from humanoid_b200 import RunningNorm

    rn = RunningNorm(934, device="cpu")
    sd = rn.state_dict()
    assert list(sd) == ["running_mean", "running_var", "count"]
    assert sd["running_mean"].shape == (1, 934) and sd["running_var"].shape == (1, 934) and sd["count"].shape == (1,)
    assert all(v.dtype == torch.float32 for v in sd.values())
    ref_like = {"running_mean": torch.full((1, 934), 0.5), "running_var": torch.full((1, 934), 2.0), "count": torch.tensor([7.0])}
    rn.load_state_dict(ref_like)
    assert float(rn.count) == 7.0 and float(rn.running_var[0, 3]) == 2.0
    rn._peers = object()  # stands in for a live PeerReduce (holds a ctypes pointer)
    buf = io.BytesIO()
    torch.save(rn, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert back._peers is None and torch.equal(back.running_mean, rn.running_mean) and back.clip == rn.clip
    assert pickle.loads(pickle.dumps(rn))._peers is None
