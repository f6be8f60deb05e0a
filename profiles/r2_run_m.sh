#!/usr/bin/env bash
# round 2 ncu --set full captures: persistent kernel at 16384 and 65536 envs, K6-fast at 4096 (the headline kernel) and at 8192
set -uo pipefail
mkdir -p gpurun_out
bash profiles/run_prof.sh r2_persist_n16384 16384 step_persist
bash profiles/run_prof.sh r2_persist_n65536 65536 step_persist
bash profiles/run_prof.sh r2_fast_n4096 4096 step_fast
bash profiles/run_prof.sh r2_fast_n8192 8192 step_fast
ls -la gpurun_out/*.ncu-rep
