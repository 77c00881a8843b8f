#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/multirank_test.log
