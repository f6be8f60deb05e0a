#!/usr/bin/env bash
# round 2, session 2: ncu --set full of the two new kernels: persistent kernel with the moments in registers (65536 envs),
# K6-fast with bookkeeping + in-step reset by role (4096 envs, clip ends only)
set -uo pipefail
mkdir -p gpurun_out
PROF_MOMENTS=1 bash profiles/run_prof.sh r2s2_persist_moments_n65536 65536 step_persist
PROF_RESET=1 bash profiles/run_prof.sh r2s2_fast_reset_n4096 4096 step_fast
ls -la gpurun_out/*r2s2*.ncu-rep
