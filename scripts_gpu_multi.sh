#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/multirank_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 1 --warmup 1 2>&1 | tail -4 | tee gpurun_out/bench_n2.log
