#!/usr/bin/env bash
# 2 GPUs: the bench as the driver launches it (peer == NCCL == single-process check, e2e floor with 2 ranks), the CUDA-IPC test
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/h_bench2.err; head -c 9000 gpurun_out/h_bench2.json
timeout 600 python -m pytest tests/test_peer_reduce_gpu.py -m gpu -q 2>&1 | tail -5
