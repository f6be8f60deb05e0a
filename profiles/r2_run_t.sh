#!/usr/bin/env bash
# in-step reset spread over the block's warps by role: parity + cost
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_pytest.log
timeout 300 python profiles/bench_step_variants2.py 4096 2>&1 | tail -12
timeout 300 python profiles/bench_env_loop.py 4096 2>&1 | tail -9
