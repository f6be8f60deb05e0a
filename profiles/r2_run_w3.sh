#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
n=16384; tag=r2s2_multi2_prefetch_T10_n16384
timeout 200 python profiles/prof_step.py $n 10 12 > gpurun_out/plain_${tag}.log 2>&1 && cat gpurun_out/plain_${tag}.log
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:step_multi2" -s 6 -c 1 -o gpurun_out/prof_${tag} python profiles/prof_step.py $n 10 12 > gpurun_out/ncu_${tag}.log 2>&1; tail -1 gpurun_out/ncu_${tag}.log
