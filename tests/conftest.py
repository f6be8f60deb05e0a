import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz fixture; ``g.inp(key)`` / ``g.out(key)`` give torch CPU tensors."""

    def __init__(self, name):
        self.name = name
        self._z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def keys(self):
        return list(self._z.keys())

    def raw(self, key):
        return torch.from_numpy(np.ascontiguousarray(self._z[key]))

    def inp(self, key):
        return self.raw("in." + key)

    def out(self, key):
        return self.raw("out." + key)

    def group(self, prefix):
        p = prefix + "."
        return {k[len(p):]: self.raw(k) for k in self._z.keys() if k.startswith(p)}


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get


STEP_CASES = ["step_aligned", "step_random", "step_T10", "step_eval_reset", "step_no_early_term"]


def _obs_groups(width):
    """(offset, vectors, floats per vector) of every column group of an obs row [self obs | T x v6 block]
    (SURVEY Appendix B); None if `width` is not an obs row."""
    for h in (1, 0):
        rest = width - (357 + h)
        if rest >= 0 and rest % 576 == 0:
            g = [(0, 1, 1)] if h else []
            g += [(h, 23, 3), (h + 69, 24, 6), (h + 213, 24, 3), (h + 285, 24, 3)]
            for t in range(rest // 576):
                o = 357 + h + 576 * t
                g += [(o, 24, 3), (o + 72, 24, 6), (o + 216, 24, 3), (o + 288, 24, 3), (o + 360, 24, 3), (o + 432, 24, 6)]
            return g
    if width % 576 == 0:  # v6 blocks only (compute_imitation_observations_v6's own output)
        g = []
        for o in range(0, width, 576):
            g += [(o, 24, 3), (o + 72, 24, 6), (o + 216, 24, 3), (o + 288, 24, 3), (o + 360, 24, 3), (o + 432, 24, 6)]
        return g
    if width % 216 == 0:  # v7 column subset [d_pos | d_vel | l_pos] per step
        return [(o, 72, 3) for o in range(0, width, 216)]
    if width in (357, 358):  # self obs only
        h = width - 357
        return ([(0, 1, 1)] if h else []) + [(h, 23, 3), (h + 69, 24, 6), (h + 213, 24, 3), (h + 285, 24, 3)]
    return None


def natural_scale(e: torch.Tensor) -> torch.Tensor:
    """Per element, the magnitude an error is to be judged against: the norm of the vector the element is a
    component of — a body's position / velocity 3-vector, a quaternion, a tangent-normal 6-vector, a joint's exp-map
    3-vector — instead of the element itself.  A rotation by a heading that is off by one ulp of atan2f moves every
    component by ~1e-7 |v|, whatever the component's own size; |v| is the scale such an error is relative to."""
    sc = e.abs().clone()
    if e.dim() >= 3 and e.shape[-1] in (3, 4):  # [.., J, 3|4]
        return e.norm(dim=-1, keepdim=True).expand_as(e).clone()
    if e.dim() == 2:
        groups = _obs_groups(e.shape[1])
        if groups is None and e.shape[1] in (69, 72):  # dof_pos / dof_vel / motion_aa: 3 per joint
            groups = [(0, e.shape[1] // 3, 3)]
        if groups is None and e.shape[1] % 196 == 0:  # AMP rows (humanoid_phc.py:469-478): [h | root rot 6 | root vel 3 |
            groups = []                               # root ang vel 3 | 19 x dof obs 6 | 19 x dof vel 3 | 4 x key pos 3] per step
            for o in range(0, e.shape[1], 196):
                groups += [(o, 1, 1), (o + 1, 1, 6), (o + 7, 2, 3), (o + 13, 19, 6), (o + 127, 19, 3), (o + 184, 4, 3)]
        if groups is not None:
            for off, n, k in groups:
                v = e[:, off : off + n * k].reshape(e.shape[0], n, k)
                sc[:, off : off + n * k] = v.norm(dim=-1, keepdim=True).expand_as(v).reshape(e.shape[0], n * k)
    return sc


def assert_close(actual, expected, rtol=1e-5, atol=1e-6, what="", scale=None, small_angle=None):
    """|a-e| <= atol + rtol*s, with a readable report of the worst entry.  s = |e| elementwise by default;
    ``scale="vec"``: s = the norm of the vector the element belongs to (``natural_scale``).  ``small_angle`` =
    dict(below, rtol, atol): for [n, 3k] exp-map rows, joints whose rotation angle |e_joint| is below ``below`` get
    their own tolerance (the reference's axis = xyz / sqrt(1 - w*w) amplifies a 1-ulp difference in w there)."""
    a = torch.as_tensor(actual).detach().cpu().double()
    e = torch.as_tensor(expected).detach().cpu().double()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    nan_a, nan_e = torch.isnan(a), torch.isnan(e)
    assert torch.equal(nan_a, nan_e), f"{what}: NaN pattern differs ({int(nan_a.sum())} vs {int(nan_e.sum())})"
    a = torch.where(nan_a, torch.zeros_like(a), a)
    e = torch.where(nan_e, torch.zeros_like(e), e)
    err = (a - e).abs()
    s = natural_scale(e) if scale == "vec" else e.abs()
    tol = atol + rtol * s
    if small_angle is not None and e.dim() == 2 and e.shape[1] % 3 == 0:
        ang = e.reshape(e.shape[0], -1, 3).norm(dim=-1, keepdim=True).expand(-1, -1, 3).reshape(e.shape)
        tol = torch.where(ang < small_angle["below"], small_angle["atol"] + small_angle["rtol"] * s, tol)
    bad = err > tol
    if bad.any():
        worst = torch.argmax(err - tol)
        idx = np.unravel_index(int(worst), a.shape) if a.dim() else ()
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{a.numel()} outside rtol={rtol} atol={atol}{' (vector-relative)' if scale else ''}; "
            f"worst at {idx}: got {a.flatten()[worst].item():.9g} want {e.flatten()[worst].item():.9g} "
            f"(abs err {err.flatten()[worst].item():.3g}, scale {s.flatten()[worst].item():.3g})"
        )


def assert_equal_exact(actual, expected, what=""):
    a = torch.as_tensor(actual).detach().cpu()
    e = torch.as_tensor(expected).detach().cpu()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    if not torch.equal(a.to(e.dtype), e):
        diff = (a.to(torch.int64) != e.to(torch.int64)).nonzero().flatten()[:8].tolist()
        raise AssertionError(f"{what}: {int((a.to(torch.int64) != e.to(torch.int64)).sum())} mismatches, first at {diff}")


def replay_episode_golden(g, new_state, update, read=lambda t: t):
    """Drive an episode-bookkeeping implementation through tests/golden/episode.npz (recorded from
    the reference's PHCPufferEnv.step) and compare after every step.  ``new_state(n, cols)`` makes the
    state dict of oracle.episode_update's layout, ``update(state, reset, terminate, rewards, raw)``
    applies one step in place, ``read`` brings a state tensor to the CPU."""
    rewards, raw = g.inp("rewards"), g.inp("reward_raw")
    reset, terminate = g.inp("reset"), g.inp("terminate")
    log = int(g.inp("log_interval"))
    infos = {int(r[0]): r[1:] for r in g.out("infos").tolist()}
    K, N = rewards.shape
    st = new_state(N, raw.shape[-1])
    count = 0
    for k in range(K):
        update(st, reset[k], terminate[k], rewards[k], raw[k])
        for key in ("terminals", "truncations", "masks", "episode_lengths"):
            assert_equal_exact(read(st[key]).to(g.out(key).dtype), g.out(key)[k], what=f"{key}[{k}]")
        # one fp32 add per env per step in the reference's order: bit-exact
        assert_equal_exact(read(st["episode_returns"]), g.out("episode_returns")[k], what=f"episode_returns[{k}]")
        if (k + 1) % log == 0:  # clean_pufferl/env.py:162-176
            s = read(st["stats"]).tolist()
            want = infos[k]
            got = [s[1] / s[0], s[2] / s[0], s[3] / s[0]] + (read(st["raw_rewards"]).double() / log).tolist()
            assert_close(torch.tensor(got), torch.tensor(want), rtol=1e-6, atol=1e-7, what=f"info[{k}]")
            count += int(s[0])
            st["stats"].zero_()
            st["raw_rewards"].zero_()
        else:
            assert_close(read(st["raw_rewards"]), g.out("raw_rewards")[k], rtol=1e-6, atol=1e-7, what=f"raw[{k}]")
        partial = 0 if (k + 1) % log == 0 else int(read(st["stats"])[0])
        assert count + partial == int(g.out("episode_count")[k]), f"episode_count[{k}]"
    assert len(infos) == K // log


def replay_env_rollout(g, env, read=lambda t: t, tol=None, dof_tol=None):
    """Drive ``env`` (oracle.OracleEnv or the CUDA shim behind the same five calls) through
    tests/golden/env_rollout.npz — a recording of the reference's own HumanoidPHC.step / reset — and
    compare every buffer after every call.  ``env`` needs: step(actions, physics) -> pd_target,
    reset(env_ids, phase), write_sim(state, dof_state, dof_force) and the reference's buffer names."""
    tol = tol or dict(rtol=1e-6, atol=1e-6)
    dof_tol = dof_tol or tol
    K = len([k for k in g.keys() if k.startswith("in.actions.")])
    for k in range(K):
        def physics(e, k=k):
            e.write_sim(g.inp(f"state.{k}"), g.inp(f"dof_state.{k}"), g.inp(f"dof_force.{k}"))

        pd = env.step(g.inp(f"actions.{k}"), physics)
        assert_equal_exact(read(pd), g.out(f"pd_target.{k}"), f"pd_target[{k}]")
        o = lambda n, k=k: g.out(f"step.{k}.{n}")  # noqa: E731
        assert_equal_exact(read(env.progress_buf), o("progress"), f"progress[{k}]")
        assert_equal_exact(read(env.reset_buf), o("reset"), f"reset[{k}]")
        assert_equal_exact(read(env._terminate_buf), o("terminate"), f"terminate[{k}]")
        assert_close(read(env.obs_buf), o("obs"), what=f"obs[{k}]", **tol)
        assert_close(read(env.rew_buf), o("rew"), what=f"rew[{k}]", **tol)
        assert_close(read(env.reward_raw), o("reward_raw"), what=f"reward_raw[{k}]", **tol)
        assert_close(read(env._amp_obs_buf).flatten(1), o("amp_obs"), what=f"amp_obs[{k}]", **tol)
        env.reset(g.inp(f"reset_indices.{k}"), g.inp(f"phase.{k}"))
        o = lambda n, k=k: g.out(f"reset.{k}.{n}")  # noqa: E731
        for name, attr in (("progress", "progress_buf"), ("reset", "reset_buf"), ("terminate", "_terminate_buf"),
                           ("motion_start_times", "_motion_start_times"), ("global_offset", "_global_offset"),
                           ("motion_start_times_offset", "_motion_start_times_offset")):  # fmt: skip
            assert_equal_exact(read(getattr(env, attr)), o(name), f"{name} after reset[{k}]")
        rows = g.inp(f"reset_indices.{k}")
        sim, root, dof = env.read_sim()
        assert_close(read(sim), o("rigid_body_state"), what=f"rigid_body_state after reset[{k}]", **tol)
        assert_close(read(root)[rows], o("root_states")[rows], what=f"root_states after reset[{k}]", **tol)
        assert_close(read(dof)[rows], o("dof_state")[rows], what=f"dof_state after reset[{k}]", **dof_tol)
        assert_close(read(env.obs_buf), o("obs"), what=f"obs after reset[{k}]", **tol)
        assert_close(read(env._amp_obs_buf).flatten(1), o("amp_obs"), what=f"amp_obs after reset[{k}]", **dof_tol)
        assert_close(read(env._amp_obs_demo_buf).flatten(1), o("amp_obs_demo"), what=f"amp_obs_demo after reset[{k}]",
                     **dof_tol)  # fmt: skip
    return K


def replay_env_reset_modes(g, env, mode, read=lambda t: t, tol=None, dof_tol=None):
    """Drive ``env`` through the ``mode`` ("Default" / "Hybrid") half of tests/golden/env_reset_modes.npz — a recording of
    the reference's own HumanoidPHC.step / reset with StateInit.Default / StateInit.Hybrid — and compare every buffer
    after every call.  ``env`` as for replay_env_rollout, with reset(env_ids, phase, default_mask)."""
    tol = tol or dict(rtol=1e-6, atol=1e-6)
    dof_tol = dof_tol or tol
    K = len([k for k in g.keys() if k.startswith(f"in.{mode}.actions.")])
    for k in range(K):
        def physics(e, k=k):
            e.write_sim(g.inp(f"{mode}.state.{k}"), g.inp(f"{mode}.dof_state.{k}"), g.inp(f"{mode}.dof_force.{k}"))

        pd = env.step(g.inp(f"{mode}.actions.{k}"), physics)
        assert_equal_exact(read(pd), g.out(f"{mode}.pd_target.{k}"), f"pd_target[{k}]")
        o = lambda n, k=k: g.out(f"{mode}.step.{k}.{n}")  # noqa: E731
        assert_equal_exact(read(env.progress_buf), o("progress"), f"progress[{k}]")
        assert_equal_exact(read(env.reset_buf), o("reset"), f"reset[{k}]")
        assert_equal_exact(read(env._terminate_buf), o("terminate"), f"terminate[{k}]")
        assert_close(read(env.obs_buf), o("obs"), what=f"obs[{k}]", **tol)
        assert_close(read(env.rew_buf), o("rew"), what=f"rew[{k}]", **tol)
        assert_close(read(env.reward_raw), o("reward_raw"), what=f"reward_raw[{k}]", **tol)
        rows = g.inp(f"{mode}.reset_indices.{k}")
        if mode == "Default":
            default_mask, phase = torch.ones(rows.shape[0], dtype=torch.bool), torch.zeros(0)
        else:
            default_mask, phase = ~g.inp(f"{mode}.ref_mask.{k}"), g.inp(f"{mode}.phase.{k}")
        env.reset(rows, phase, default_mask)
        o = lambda n, k=k: g.out(f"{mode}.reset.{k}.{n}")  # noqa: E731
        for name, attr in (("progress", "progress_buf"), ("reset", "reset_buf"), ("terminate", "_terminate_buf"),
                           ("motion_start_times", "_motion_start_times"), ("global_offset", "_global_offset"),
                           ("motion_start_times_offset", "_motion_start_times_offset")):  # fmt: skip
            assert_equal_exact(read(getattr(env, attr)), o(name), f"{name} after reset[{k}]")
        sim, root, dof = env.read_sim()
        assert_close(read(sim), o("rigid_body_state"), what=f"rigid_body_state after reset[{k}]", **tol)
        assert_close(read(root)[rows], o("root_states")[rows], what=f"root_states after reset[{k}]", **tol)
        assert_close(read(dof)[rows], o("dof_state")[rows], what=f"dof_state after reset[{k}]", **dof_tol)
        assert_close(read(env.obs_buf), o("obs"), what=f"obs after reset[{k}]", **tol)
    return K
