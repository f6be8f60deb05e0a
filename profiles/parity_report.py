"""Parity margins of the CUDA path vs the CPU oracle (run on the GPU box).

For each workload and output group: max abs error, the max error RELATIVE TO THE NORM OF THE VECTOR the entry is a
component of (a body's 3-vector, a tangent-normal 6-vector, a quaternion: tests/conftest.py natural_scale — the
natural scale of an error that comes from rotating the vector by a heading that is 1 ulp of atan2f off), the
per-element relative error for reference, and the worst ratio err / (atol + rtol * |vector|) with the test tolerances
(rtol 1e-5, atol 1e-6) — a ratio below 1 passes; how far below is the margin.  Integer / flag outputs are compared
exactly.

    python profiles/parity_report.py [out.md]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, synth  # noqa: E402
from oracle import phc_oracle as O  # noqa: E402

RTOL, ATOL = 1e-5, 1e-6
from conftest import natural_scale  # noqa: E402
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
dev = "cuda"


def q(lib_data, ids, times, offset):
    return O.OracleMotionLib(lib_data).get_motion_state(ids, times, offset)


CASES = {
    "config1 N=256 M=16": dict(num_envs=256, num_motions=16, seed=101, max_progress=30),
    "config2 N=4096 ids=arange": dict(num_envs=4096, num_motions=4096, seed=102, max_frames=120, max_progress=40),
    "config4 random ids, mixed fps, unaligned": dict(num_envs=2048, num_motions=96, seed=103, ids="random", aligned=False,
                                                     fps_choices=(30, 60, 120), min_frames=60, max_frames=900, max_progress=40),
    "i.i.d. random rotations": dict(num_envs=1024, num_motions=64, seed=104, rot_regime="random", max_progress=40),
}  # fmt: skip


def stats(name, got, want, layout_of=None):
    """``layout_of``: the full tensor whose column layout `want` is a slice of (a slice alone has lost its layout)."""
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    err = (got - want).abs()
    scale = natural_scale(want) if layout_of is None else layout_of
    vrel = err / scale.clamp_min(1e-30)
    rel = err / want.abs().clamp_min(1e-30)
    ratio = err / (ATOL + RTOL * scale)
    sig = want.abs() > 1e-3
    vsig = scale > 1e-2
    print(f"| {name} | {err.max():.3g} | {vrel[vsig].max() if vsig.any() else 0:.3g} | {rel[sig].max() if sig.any() else 0:.3g} | "
          f"{ratio.max():.3f} |", file=out)
    return float(ratio.max())


print(f"# Parity margins vs the oracle (rtol {RTOL} of the vector norm, atol {ATOL})\n", file=out)
worst = 0.0
for name, kw in CASES.items():
    lib_data, clock, state = synth.make_case(query=q, device="cpu", **kw)
    prog = clock.progress_buf.clone()
    want = O.step(O.OracleMotionLib(lib_data), state, prog, clock.motion_start_times, clock.motion_start_times_offset,
                  clock.global_offset, clock.sampled_motion_ids, torch.full((24,), 0.25), synth.SIM_DT)  # fmt: skip
    lib = MotionLib(lib_data, device=dev)
    env = HumanoidPHC(lib, state.shape[0], device=dev)
    env.set_sim_state(state.to(dev))
    env.set_clock(clock.to(dev))
    env.step()
    torch.cuda.synchronize()
    print(f"## {name}\n\n| output | max abs err | max err / |vector| (|vector| > 1e-2) | max per-element rel err (|ref| > 1e-3) | "
          f"worst err/tol |\n|---|---|---|---|---|", file=out)
    obs = env.obs_buf
    sc = natural_scale(want[0].double())
    grp = {"self: height": (0, 1), "self: body pos (local)": (1, 70), "self: body rot (tan-norm)": (70, 214),
           "self: body vel": (214, 286), "self: body ang vel": (286, 358)}
    for k, (a, b) in grp.items():
        worst = max(worst, stats(k, obs[:, a:b], want[0][:, a:b], sc[:, a:b]))
    blk = {"d_pos": (0, 72), "d_rot": (72, 216), "d_vel": (216, 288), "d_ang": (288, 360), "l_pos": (360, 432), "l_rot": (432, 576)}
    for k, (a, b) in blk.items():
        worst = max(worst, stats(f"task obs {k}", obs[:, 358 + a : 358 + b], want[0][:, 358 + a : 358 + b], sc[:, 358 + a : 358 + b]))
    worst = max(worst, stats("reward", env.rew_buf, want[1]))
    for i, k in enumerate(("r_pos", "r_rot", "r_vel", "r_ang_vel")):
        worst = max(worst, stats(f"reward_raw {k}", env.reward_raw[:, i], want[2][:, i]))
    flags = (int((env.reset_buf.cpu() != want[3]).sum()), int((env._terminate_buf.cpu() != want[4]).sum()),
             int((env.progress_buf.cpu() != prog).sum()))  # fmt: skip
    print(f"\nreset / terminated / progress mismatches: {flags[0]} / {flags[1]} / {flags[2]} of {state.shape[0]} "
          f"(terminated fraction {float(want[4].float().mean()):.3f})\n", file=out)
    t = synth.reward_time(clock, extra_steps=1)
    ref = O.OracleMotionLib(lib_data).get_motion_state(clock.sampled_motion_ids, t, clock.global_offset)
    got = lib.get_motion_state(clock.sampled_motion_ids.to(dev), t.to(dev), clock.global_offset.to(dev), with_frame_info=True)
    print("| get_motion_state output | max abs err | max err / |vector| | max per-element rel err | worst err/tol |\n|---|---|---|---|---|", file=out)
    for k in ("rg_pos", "rb_rot", "body_vel", "body_ang_vel", "dof_vel", "dof_pos"):
        stats(k, got[k], ref[k])
    # dof_pos by joint angle: the 1e-5 bar holds wherever the joint's rotation is not tiny (tests/util_gpu.py DOF_TOL)
    dp_g, dp_w = got["dof_pos"].cpu().double().reshape(-1, 23, 3), ref["dof_pos"].double().reshape(-1, 23, 3)
    ang = dp_w.norm(dim=-1)
    verr = (dp_g - dp_w).abs().amax(dim=-1) / ang.clamp_min(1e-30)
    big = ang >= 0.283
    print(f"\ndof_pos, joints with angle >= 0.283 rad (|w| <= 0.99): {int(big.sum())} of {big.numel()}, max err / angle "
          f"{float(verr[big].max()) if big.any() else 0:.3g}; smaller angles: max err / angle {float(verr[~big & (ang > 1e-4)].max()) if (~big).any() else 0:.3g}", file=out)
    print(f"\nframe_idx0 / frame_idx1 mismatches: {int((got['frame_idx0'].cpu() != ref['frame_idx0']).sum())} / "
          f"{int((got['frame_idx1'].cpu() != ref['frame_idx1']).sum())}; blend bit-exact: "
          f"{bool(torch.equal(got['blend'].cpu(), ref['blend']))}\n", file=out)
print(f"worst err/tol over all step outputs: {worst:.3f}", file=out)
