#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/d_pytest.log

timeout 300 python profiles/bench_step_variants2.py > gpurun_out/d_variants2.log 2>&1; echo "variants2 rc=$?"; tail -20 gpurun_out/d_variants2.log
timeout 300 python profiles/bench_env_loop.py > gpurun_out/d_envloop.log 2>&1; echo "envloop rc=$?"; tail -12 gpurun_out/d_envloop.log
