#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/f_pytest.log
timeout 300 python profiles/bench_step_variants2.py > gpurun_out/f_variants2.log 2>&1; echo "variants2 rc=$?"; tail -14 gpurun_out/f_variants2.log
timeout 300 python profiles/bench_env_loop.py > gpurun_out/f_envloop.log 2>&1; echo "envloop rc=$?"; tail -9 gpurun_out/f_envloop.log
timeout 600 python profiles/bench_persist.py > gpurun_out/f_persist.log 2>&1; echo "persist rc=$?"; tail -9 gpurun_out/f_persist.log
