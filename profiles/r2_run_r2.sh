#!/usr/bin/env bash
# host pipeline timeline, staged path
set -uo pipefail
mkdir -p gpurun_out
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 20 --no-long-run"
for c in 2 3 4 6; do
  echo "== staged chunks $c"
  PHC_HOST_TRACE=1 PHC_HOST_PATH=staged timeout 300 python bench.py $B --e2e-chunks $c > gpurun_out/r_e2e_s$c.json 2> gpurun_out/r_e2e_s$c.err
  grep -A8 "phc_host trace" gpurun_out/r_e2e_s$c.err | head -8
  python -c "
import json; d=json.load(open('gpurun_out/r_e2e_s$c.json'))['e2e']; print('e2e us', round(d['us_per_step'],1), 'floor', round(d['floor']['us_per_step'],1), 'd2h only', round(d['floor']['d2h_only_us'],1))"
done
