"""K6-multi (one thread group per block) against K6-multi2 (two query groups per block) on config 5 (T = 10).

    python profiles/bench_multi_groups.py [num_envs ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = [int(a) for a in sys.argv[1:]] or [2048, 16384]
print("| envs / GPU | groups | us / step | of measured HBM peak |\n|---|---|---|---|")
for n in sizes:
    for g in (1, 2):
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); from humanoid_b200 import _cabi; "
                f"_cabi.load().phc_set_option(_cabi.OPT_MULTI_GROUPS, {g}); "
                f"sys.argv = ['bench.py', '--time-steps', '10', '--num-envs', '{n}', '--no-cpu-baseline', '--steps', '128', '--e2e-steps', '2']; "
                "import bench; bench.main()")
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT).stdout
        d = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
        print(f"| {n} | {g} | {d['ms_per_step'] * 1e3:.2f} | {d['roofline']['frac']:.3f} |", flush=True)
