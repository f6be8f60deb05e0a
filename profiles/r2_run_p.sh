#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
q() { python -c "
import json,sys
d=json.load(open('$1')); print('$2', 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'])"; }
B="--no-extra-configs --no-cpu-baseline --no-torch-cuda-baseline --steps 64 --no-long-run"
for p in 0 1; do PHC_STEP_PERSIST=$p timeout 300 python bench.py $B --workload config4 --num-envs 16384 > gpurun_out/p_c4_16384_p$p.json 2>/dev/null; q gpurun_out/p_c4_16384_p$p.json "config4 16384 persist=$p"; done
for p in 0 1; do PHC_STEP_PERSIST=$p timeout 300 python bench.py $B --workload config4 --num-envs 65536 > gpurun_out/p_c4_65536_p$p.json 2>/dev/null; q gpurun_out/p_c4_65536_p$p.json "config4 65536 persist=$p"; done
for n in 10240 12288 14336; do for p in 0 2; do PHC_STEP_PERSIST=$p timeout 300 python bench.py $B --num-envs $n > gpurun_out/p_c2_${n}_p$p.json 2>/dev/null; q gpurun_out/p_c2_${n}_p$p.json "config2 $n persist=$p"; done; done
