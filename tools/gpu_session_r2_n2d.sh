#!/bin/bash
# 2-GPU session D (closing state): world-size-2 NCCL tests, a short N=2 bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_n2d_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29717 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_n2d_bench.json 2> gpurun_out/r2_n2d_bench.err; tail -c 900 gpurun_out/r2_n2d_bench.json; tail -2 gpurun_out/r2_n2d_bench.err
