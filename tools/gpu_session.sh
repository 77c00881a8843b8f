#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder_ops.py tests/test_gpu_encoder.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/prof_chunk.log 2>&1
