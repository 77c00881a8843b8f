#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_infonce.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/infonce_test.log
timeout 600 python -m pytest tests/test_gpu_encoder.py -q -m gpu 2>&1 | tail -80 | tee gpurun_out/encoder_test.log
