#!/bin/bash
# round-2 session B: first hardware run of the 4-CTA-cluster GEMM (B multicast), the fused SwiGLU-backward epilogue and the
# in-epilogue RoPE; each new kernel family in its own process (a trapped kernel poisons the context), then A/B timings + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -k "quad" 2>&1 | tail -6 | tee gpurun_out/r2b_quad.log
if ! grep -q " passed" gpurun_out/r2b_quad.log || grep -q "failed\|error" gpurun_out/r2b_quad.log; then export CX_NO_QUAD=1; echo "QUAD FAILED -> CX_NO_QUAD=1" | tee -a gpurun_out/r2b_quad.log; fi
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -k "swiglu_bwd" 2>&1 | tail -6 | tee gpurun_out/r2b_swiglu_bwd.log
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -k "rope" 2>&1 | tail -6 | tee gpurun_out/r2b_rope.log
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2b_tests.log
timeout 300 python tools/bench_kernels.py > gpurun_out/r2b_kernels.log 2>&1; grep "cluster" gpurun_out/r2b_kernels.log | cut -c1-330
timeout 120 python tools/bench_qkv_rope.py 2>&1 | tail -4 | tee gpurun_out/r2b_qkv_rope.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-baseline --no-selfcheck > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2500 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2b_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches_chunk.csv 0.5 > gpurun_out/r2b_launches_chunk_summary.txt; head -16 gpurun_out/r2b_launches_chunk_summary.txt
