#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.load(open('$1')); print('$2', 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'])"; }
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 32 --no-long-run --e2e-steps 1 --time-steps 10"
for n in 2048 8192 16384; do for pf in 1 2 1 2; do
PHC_MULTI_PREFETCH=$pf timeout 300 python bench.py $B --num-envs $n > gpurun_out/mp.json 2>/dev/null; q gpurun_out/mp.json "T10 $n prefetch=$pf"
done; done
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 64 --no-long-run --e2e-steps 1"
for n in 10240 16384 32768 65536; do for pf in 0 1 0 1; do
PHC_PERSIST_PREFETCH=$pf timeout 300 python bench.py $B --num-envs $n > gpurun_out/mp.json 2>/dev/null; q gpurun_out/mp.json "T1 $n persist_prefetch=$pf"
done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_T or T10 or config5 or trajectory or persist" 2>&1 | tail -2
