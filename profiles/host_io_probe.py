"""Where does the end-to-end time go?  Times phc_step_fused with inputs and/or outputs placed in
pinned host memory (mapped into the device address space) instead of HBM."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from humanoid_b200 import HumanoidPHC, MotionLib, _cabi, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
state = synth.make_sim_state(ref, seed=1236)
env = HumanoidPHC(lib, N, device=dev)
env.set_sim_state(state)
env.set_clock(clock)
env.step()
capi = _cabi.load()
h_state = state.cpu().pin_memory()
h_obs = torch.empty(N, 934).pin_memory()
h_small = {k: v.cpu().pin_memory() for k, v in dict(rew=env.rew_buf, raw=env.reward_raw, reset=env.reset_buf, term=env._terminate_buf).items()}


def args_for(state_t, obs_t, small_host, flags):
    a = env._build_step_args(True)
    st = state_t
    V = _cabi.PhcView
    a.body = _cabi.PhcBodyState(V(st.data_ptr(), 312, 13), V(st.data_ptr() + 12, 312, 13), V(st.data_ptr() + 28, 312, 13),
                                V(st.data_ptr() + 40, 312, 13), 24)
    a.obs_buf = obs_t.data_ptr()
    if small_host:
        a.rew_buf, a.reward_raw = h_small["rew"].data_ptr(), h_small["raw"].data_ptr()
        a.reset_buf, a.terminate_buf = h_small["reset"].data_ptr(), h_small["term"].data_ptr()
    a.flags = flags
    return a


def timeit(a, reps=30):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        capi.phc_step_fused(lib.handle, C.byref(a), N, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        capi.phc_step_fused(lib.handle, C.byref(a), N, s)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


d_state = env._rigid_body_state_reshaped
print(f"N={N}")
print(f"all device            : {timeit(args_for(d_state, env.obs_buf, False, 0)):8.1f} us")
print(f"sim state from host   : {timeit(args_for(h_state, env.obs_buf, False, 1)):8.1f} us   (5.1 MB over PCIe -> >= 98 us)")
print(f"obs to host           : {timeit(args_for(d_state, h_obs, False, 1)):8.1f} us   (15.3 MB over PCIe -> >= 273 us)")
print(f"obs + scalars to host : {timeit(args_for(d_state, h_obs, True, 1)):8.1f} us")
print(f"both                  : {timeit(args_for(h_state, h_obs, True, 1)):8.1f} us   (full duplex -> >= 295 us)")
