#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.load(open('$1')); print('$2', 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'])"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_T or T10 or config5 or trajectory" 2>&1 | tail -2
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 32 --no-long-run --e2e-steps 1 --time-steps 10"
for n in 2048 4096 8192 16384; do for r in 1 2; do
timeout 300 python bench.py $B --num-envs $n > gpurun_out/mp.json 2>/dev/null; q gpurun_out/mp.json "T10 $n distributed refill"
done; done
