"""Motion-library build: phc_motion_build (three kernels over all frames) timed on the device, and the
oracle's per-clip CPU restatement of the reference's load_motions on a bounded sample of the same clips.

    python profiles/bench_build.py [num_clips]        (default 4096: the reference's one-clip-per-env regime)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from humanoid_b200 import _cabi  # noqa: E402
from humanoid_b200.motion_build import build_motion_tensors  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
PARENTS = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]
nf = rng.integers(60, 301, size=M)
F = int(nf.sum())
quat = rng.normal(size=(F, 24, 4))
quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
trans = rng.normal(size=(F, 3))
aa = rng.normal(size=(F, 72))
lt = (rng.normal(size=(M, 24, 3)) * 0.2).astype(np.float32)
fps = np.full(M, 30)
u = rng.random(M)

dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
_cabi.load()
host_s = {}
for pin in (False, True, False, True):  # host arrays -> library tensors in HBM, incl. the H2D copy of the fp64 inputs
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = build_motion_tensors(quat, trans, aa, nf, fps, PARENTS, lt, heading_u=u, device=dev, pin=pin)
    torch.cuda.synchronize()
    host_s["pinned_staging" if pin else "pageable"] = time.perf_counter() - t0

# device-only: inputs resident, time the three launches
d = {k: torch.from_numpy(v).to(dev) for k, v in dict(quat=quat, trans=trans, aa=aa).items()}
starts = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(np.int64)
d_lt, d_nf, d_st = torch.from_numpy(lt).to(dev), torch.from_numpy(nf.astype(np.int64)).to(dev), torch.from_numpy(starts).to(dev)
d_fps = torch.from_numpy(fps.astype(np.float64)).to(dev)
from humanoid_b200.motion_build import gaussian_taps, heading_half_angle  # noqa: E402
import ctypes as C  # noqa: E402

d_head = torch.from_numpy(heading_half_angle(u)).to(dev)
scratch = torch.empty(F * 108, dtype=torch.float64, device=dev)
parents = (C.c_int32 * 24)(*PARENTS)
taps = (C.c_double * 17)(*gaussian_taps().tolist())
args = _cabi.PhcBuildArgs(
    d["quat"].data_ptr(), d["trans"].data_ptr(), d["aa"].data_ptr(), d_lt.data_ptr(), d_nf.data_ptr(), d_st.data_ptr(),
    d_fps.data_ptr(), d_head.data_ptr(), parents, taps, F, M, out["gts"].data_ptr(), out["grs"].data_ptr(),
    out["lrs"].data_ptr(), out["gvs"].data_ptr(), out["gavs"].data_ptr(), out["dvs"].data_ptr(),
    out["motion_aa"].data_ptr(), scratch.data_ptr())  # fmt: skip
lib = _cabi.load()
s = torch.cuda.current_stream(dev)
for _ in range(3):
    _cabi.check(lib.phc_motion_build(C.byref(args), s.cuda_stream), "build")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    lib.phc_motion_build(C.byref(args), s.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# algorithmic bytes per frame: fp64 inputs 768 + 24 + 576; fp32 outputs 288+384+384+288+288+276+288; scratch (fp64 + fp32) w+r
bytes_per_frame = (768 + 24 + 576) + (288 + 384 + 384 + 288 + 288 + 276 + 288) + 2 * (576 + 288)
res = {"clips": M, "frames": F, "device_ms": ms, "frames_per_s": F / ms * 1e3, "GBps": bytes_per_frame * F / ms / 1e6,
       "bytes_per_frame": bytes_per_frame, "host_call_s_incl_h2d": host_s}  # fmt: skip

# CPU: the oracle (per clip, like the reference's worker) on a bounded sample
from oracle import build_oracle as B  # noqa: E402

torch.set_num_threads(1)
sample = 32
st = np.concatenate([[0], np.cumsum(nf)])
t0 = time.perf_counter()
for m in range(sample):
    sl = slice(int(st[m]), int(st[m + 1]))
    B.build_motion_library(quat[sl], trans[sl], aa[sl], [nf[m]], [30], PARENTS, lt[m:m + 1], np.zeros((1, 17)),
                           np.zeros((1, 10)), heading_u=[u[m]])  # fmt: skip
cpu_s = time.perf_counter() - t0
res["cpu_oracle"] = {"clips": sample, "frames": int(st[sample]), "s": cpu_s, "frames_per_s": int(st[sample]) / cpu_s,
                     "threads": 1, "extrapolated_s_for_all_clips": cpu_s * F / int(st[sample])}  # fmt: skip
print(json.dumps(res))
