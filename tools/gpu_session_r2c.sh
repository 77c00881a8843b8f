#!/bin/bash
# round-2 session C: double-buffered SwiGLU-backward epilogue, dK rotary in the attention backward epilogue, pair-mode default,
# real-dimension ViT + LiT tests; then micro-timings, the bench line and the launch list of one chunk
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -k "swiglu_bwd" 2>&1 | tail -4 | tee gpurun_out/r2c_swiglu_bwd.log
timeout 300 python -m pytest tests/test_gpu_encoder_ops.py -q -m gpu -x -k "attention" 2>&1 | tail -4 | tee gpurun_out/r2c_attn.log
timeout 600 python -m pytest tests/test_gpu_vit.py -q -m gpu -x -s 2>&1 | tail -6 | tee gpurun_out/r2c_vit.log
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2c_tests.log
timeout 300 python tools/bench_kernels.py > gpurun_out/r2c_kernels.log 2>&1; grep "swiglu\|infonce" gpurun_out/r2c_kernels.log | cut -c1-300
timeout 120 python tools/bench_attn.py 2>&1 | grep "bert\|vit" | cut -c1-330 | tee gpurun_out/r2c_attn_bench.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-baseline --no-selfcheck > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2c_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c_launches_chunk.csv 0.5 > gpurun_out/r2c_launches_chunk_summary.txt; head -20 gpurun_out/r2c_launches_chunk_summary.txt
