#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/tests_all_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 1200 python bench.py --steps 1 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_pair.log
