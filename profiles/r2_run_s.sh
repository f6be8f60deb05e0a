#!/usr/bin/env bash
# tuned host schedule: tests + e2e with the tuner's table
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline" 2>&1 | tail -3
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 20 --no-long-run"
for i in 1 2; do
  PHC_HOST_TRACE=1 timeout 300 python bench.py $B > gpurun_out/s_e2e_$i.json 2> gpurun_out/s_e2e_$i.err
  grep "phc_host tune" gpurun_out/s_e2e_$i.err
  python -c "
import json; d=json.load(open('gpurun_out/s_e2e_$i.json'))['e2e']; print('e2e us', round(d['us_per_step'],1), 'chunks', d['chunks'], d['output_path'][:7], 'floor', round(d['floor']['us_per_step'],1), 'd2h only', round(d['floor']['d2h_only_us'],1))"
done
