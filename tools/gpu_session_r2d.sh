#!/bin/bash
# round-2 session D: pooler tests, narrow LayerNorm backward, polynomial-exp A/B of the attention forward, ncu --set full captures
# (attention forward / backward with source, LayerNorm kernels, the fused SwiGLU-backward GEMM), bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_poolers.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2d_poolers.log
timeout 300 python -m pytest tests/test_gpu_encoder_ops.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r2d_encoder_ops.log
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r2d_tests.log
for p in 0 4 8; do CX_ATTN_POLY=$p timeout 120 python tools/bench_attn.py 2>&1 | grep "8 3" | cut -c1-200 | sed "s/^/poly=$p /" | tee -a gpurun_out/r2d_attn_poly.log; done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"attn_(fwd4|bwd3)" -s 6 -c 2 -o gpurun_out/r02d_attn python tools/bench_attn.py > gpurun_out/r2d_ncu_attn.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"add_layernorm_(bwd_narrow|fwd)" -s 60 -c 2 -o gpurun_out/r02d_layernorm python tools/profile_chunk.py 1 > gpurun_out/r2d_ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gemm_kernel<256, 0, 1, 4" -c 1 -o gpurun_out/r02d_gemm_swiglu_bwd python tools/profile_chunk.py 1 > gpurun_out/r2d_ncu_swiglu.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-baseline --no-selfcheck > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1200 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2d_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2d_launches_chunk.csv 0.5 > gpurun_out/r2d_launches_chunk_summary.txt; head -14 gpurun_out/r2d_launches_chunk_summary.txt
ls -la gpurun_out/*.ncu-rep
