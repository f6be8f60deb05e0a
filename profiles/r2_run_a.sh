#!/usr/bin/env bash
# round 2, GPU call A: parity suite, then the bench the way the driver runs it, then the default bench
set -uo pipefail
mkdir -p gpurun_out
nproc; nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/a_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/a_bench20.json 2> gpurun_out/a_bench20.err; echo "bench20 rc=$?"; tail -5 gpurun_out/a_bench20.err; head -c 6000 gpurun_out/a_bench20.json
timeout 600 python bench.py > gpurun_out/a_bench512.json 2> gpurun_out/a_bench512.err; echo "bench512 rc=$?"; tail -5 gpurun_out/a_bench512.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_benchref.json 2> gpurun_out/a_benchref.err; echo "ref rc=$?"; head -c 1500 gpurun_out/a_benchref.json
