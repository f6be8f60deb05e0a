#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/o_pytest.log
timeout 300 python profiles/bench_step_variants2.py > gpurun_out/o_variants2.log 2>&1; echo "variants2 rc=$?"; tail -14 gpurun_out/o_variants2.log
timeout 600 python profiles/bench_rollout_moments.py > gpurun_out/o_rollout.log 2>&1; echo "rollout rc=$?"; tail -7 gpurun_out/o_rollout.log
