#!/usr/bin/env bash
# round 2, session 2: the driver's sequence on one GPU
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/x_bench20.json 2> gpurun_out/x_bench20.err; echo "bench20 rc=$?"; tail -3 gpurun_out/x_bench20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/x_benchref.json 2> gpurun_out/x_benchref.err; echo "ref rc=$?"
