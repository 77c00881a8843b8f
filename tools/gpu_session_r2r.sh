#!/bin/bash
# session R: 2048-token attention case, compute-sanitizer memcheck over one ragged attention forward + backward
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encoder_ops.py -q -x -k "attention" 2>&1 | tail -3 | tee gpurun_out/r2r_tests.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_attn.py > gpurun_out/r2r_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/r2r_memcheck.log
