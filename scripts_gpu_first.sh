#!/bin/bash
# first GPU bring-up: unit tests with a hard timeout each
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/gemm_test.log
timeout 600 python -m pytest tests/test_gpu_infonce.py -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/infonce_test.log
