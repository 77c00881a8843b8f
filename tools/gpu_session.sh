#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 300 python tools/bench_kernels.py 2>&1 | tail -3 | tee gpurun_out/bench_kernels.log
