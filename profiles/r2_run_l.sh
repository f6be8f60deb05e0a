#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/l_pytest.log
timeout 600 python profiles/bench_persist.py > gpurun_out/l_persist.log 2>&1; echo "persist rc=$?"; tail -8 gpurun_out/l_persist.log
