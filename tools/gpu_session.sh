#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder_ops.py tests/test_gpu_encoder.py tests/test_gpu_vit.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/tests.log
