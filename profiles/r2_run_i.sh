#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python profiles/parity_report.py gpurun_out/i_parity.md; echo "parity rc=$?"; tail -3 gpurun_out/i_parity.md
bash profiles/r2_traffic.sh
