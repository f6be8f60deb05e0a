#!/usr/bin/env bash
# round 2, GPU call B: parity suite (new: in-step reset, flags, res_action, consistency), step-variant timings
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/b_pytest.log
timeout 300 python bench.py --no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(d['ms_per_step'], d['roofline']['frac'])"
