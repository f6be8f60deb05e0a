#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persist or moments" 2>&1 | tail -4
timeout 900 python profiles/bench_moments_sizes.py 2>&1 | tail -8
