#!/usr/bin/env bash
# round 2, session 2, final binary: ncu --set full of the headline kernel (K6-fast, 4096 envs), of config 4 and of the T = 10 kernel (16384 envs)
set -uo pipefail
mkdir -p gpurun_out
bash profiles/run_prof.sh r2s2_fast_n4096 4096 step_fast
n=4096; tag=r2s2_fast_config4_n4096
timeout 200 python profiles/prof_step.py $n 1 40 config4 > gpurun_out/plain_${tag}.log 2>&1 && cat gpurun_out/plain_${tag}.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:step_fast" -s 20 -c 1 -o gpurun_out/prof_${tag} python profiles/prof_step.py $n 1 40 config4 > gpurun_out/ncu_${tag}.log 2>&1; tail -1 gpurun_out/ncu_${tag}.log
n=16384; tag=r2s2_multi2_T10_n16384
timeout 200 python profiles/prof_step.py $n 10 12 > gpurun_out/plain_${tag}.log 2>&1 && cat gpurun_out/plain_${tag}.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:step_multi2" -s 6 -c 1 -o gpurun_out/prof_${tag} python profiles/prof_step.py $n 10 12 > gpurun_out/ncu_${tag}.log 2>&1; tail -1 gpurun_out/ncu_${tag}.log
ls -la gpurun_out/prof_r2s2_fast_n4096.ncu-rep gpurun_out/prof_r2s2_fast_config4_n4096.ncu-rep gpurun_out/prof_r2s2_multi2_T10_n16384.ncu-rep
