"""Each piece of PHCPufferEnv.step alone in a 64-launch CUDA graph (N = 4096): where the env-side loop spends its time.

    python profiles/bench_env_breakdown.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from humanoid_b200 import HumanoidPHC, MotionLib, PHCPufferEnv, synth
N=4096; dev=torch.device("cuda",0)
lib_data = synth.make_motion_lib(N, 60, 300, (30,), seed=1234, device=dev)
lib = MotionLib(lib_data, device=dev)
clock = synth.make_clock(lib_data, N, seed=1235, max_progress=30)
env = HumanoidPHC(lib, N, device=dev, use_power_reward=True, use_amp_obs=True)
ref = lib.get_motion_state(clock.sampled_motion_ids, synth.reward_time(clock, extra_steps=1), clock.global_offset)
env.set_sim_state(synth.make_sim_state(ref, seed=1236)); env.set_clock(clock)
penv = PHCPufferEnv(env, log_interval=1<<30, use_amp_obs=True)
actions = torch.rand(N,69,device=dev)*2.4-1.2; phase=torch.rand(N,device=dev)
state0 = env._rigid_body_state_reshaped.clone()
K=64
def timed(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream()); g=torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(K): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    best=1e9
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record(s); g.replay(); e1.record(s)
        torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1)/K*1e3)
    print(f"{name}: {best:.2f} us", flush=True)
mask_all = torch.ones(N, dtype=torch.bool, device=dev)
timed(lambda: env._rigid_body_state_reshaped.copy_(state0), "state copy (physics stand-in)")
timed(lambda: torch.clamp(actions,-1,1,out=penv.actions), "clamp actions")
timed(lambda: env.post_physics_step(True), "fused step + power reward")
timed(lambda: env._amp_step(roll=True), "amp step (roll + slot 0)")
timed(lambda: penv.rewards.clone(), "reward clone")
timed(lambda: penv.update_episodes(), "episode update")
timed(lambda: env.reset_buf.clone(), "reset_buf clone")
env.use_amp_obs=False
timed(lambda: env.reset_done(phase), "reset_done without amp (reset kernel + masked obs pass)")
env.use_amp_obs=True
m = env.reset_buf.clone()
timed(lambda: env._init_amp_obs_masked(m), "amp init for flagged envs")
print("flagged fraction", float(env.reset_buf.float().mean()))
# with a realistic share of flagged envs (the synthetic state terminates ~24 % per step)
env.set_clock(clock)
env._rigid_body_state_reshaped.copy_(state0)
env.post_physics_step(True)
m = env.reset_buf.clone()
print("flagged fraction", float(m.float().mean()))
timed(lambda: env._init_amp_obs_masked(m), "amp init, 24 % flagged")
env.use_amp_obs = False
timed(lambda: env._reset_masked(m, phase), "reset kernel + masked obs pass, 24 % flagged")
env.use_amp_obs = True
timed(lambda: env._reset_masked(m, phase), "reset incl. amp init, 24 % flagged")
