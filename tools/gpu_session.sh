#!/bin/bash
# full GPU parity suite + attention micro-timings (one gpurun call)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests.log; tail -12 gpurun_out/tests.log
timeout 90 python tools/bench_attn.py 6,2 2>&1 | grep "bert\|vit"
timeout 120 python tools/trace_attn.py > gpurun_out/trace6.log 2>&1; grep -c ready gpurun_out/trace6.log
