#!/bin/bash
# First hardware run of the experimental attention kernels (one gpurun call): parity per kernel in its own process (a trapped
# kernel poisons the CUDA context), then CUDA-event timings next to the default generation.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for k in two-threads persistent transposed; do
  CX_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_encoder_ops.py -q -m gpu -x -k "attention and experimental and $k" 2>&1 | tail -4 | tee gpurun_out/experimental_$k.log
done
for m in 6,2 8,2 9,2 6,3; do
  timeout 90 python tools/bench_attn.py $m 2>&1 | grep "bert\|vit"
done
