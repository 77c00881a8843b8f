#!/bin/bash
# 2-GPU session (gpurun --gpus 2): the world-size-2 NCCL parity tests and a short N=2 bench (comm_ms, bucketed gradient reduction)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2_n2_multirank_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 2 --warmup 2 --no-selfcheck > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err; tail -c 2500 gpurun_out/r2_n2_bench.json; tail -3 gpurun_out/r2_n2_bench.err
