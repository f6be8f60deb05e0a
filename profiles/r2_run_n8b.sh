#!/usr/bin/env bash
# 8 GPUs, session 2: the bench as the driver launches it (tuned host schedule, persistent kernel with moments in the config 3 rollout)
set -uo pipefail
mkdir -p gpurun_out
PHC_HOST_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8b_auto.json 2> gpurun_out/n8b_auto.err; echo "rc=$?"
grep "phc_host tune" gpurun_out/n8b_auto.err | head -12
python - <<PY
import json
d=json.load(open("gpurun_out/n8b_auto.json"))
e=d["e2e"]; print("value %.3g"%d["value"], "us/step %.2f"%(d["ms_per_step"]*1e3), "e2e %.3g"%e["value"], "e2e us %.0f"%e["us_per_step"], e["output_path"][:8], "chunks", e["chunks"], "floor us %.0f"%e["floor"]["us_per_step"], "e2e/floor %.2f"%e["floor"]["e2e_over_floor"])
print({k:v for k,v in d["rms"].items() if k.startswith(("peer","ranks","exchange"))})
for k,v in d.get("configs",{}).items(): print(k, "%.2f us"%v["us_per_step"], "frac %.3f"%v["frac"], ("rollout %.2f us/step"%v["rollout"]["us_per_step"]) if "rollout" in v else "")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8b_ref.json 2>/dev/null; echo "ref rc=$?"; head -c 300 gpurun_out/n8b_ref.json
