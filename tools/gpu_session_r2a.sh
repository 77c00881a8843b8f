#!/bin/bash
# round-2 session A (one gpurun call): full GPU parity suite with the new BASELINE-shape tests, attention timings next to FA2,
# GEMM / InfoNCE micro-timings, the bench line (with the reference-on-GPU leg), and the two ncu launch lists
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s 2>&1 | tail -15 | tee gpurun_out/r2a_tests.log
timeout 120 python tools/bench_attn.py 8,3 2>&1 | grep "bert\|vit" | tee gpurun_out/r2a_attn.log
timeout 200 python tools/bench_kernels.py > gpurun_out/r2a_kernels.log 2>&1; tail -5 gpurun_out/r2a_kernels.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 4000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2a_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2a_launches_chunk.csv 0.5 > gpurun_out/r2a_launches_chunk_summary.txt; head -16 gpurun_out/r2a_launches_chunk_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_ref_gpu.csv python tools/profile_ref_gpu.py 2 > gpurun_out/r2a_prof_ref.log 2>&1
python tools/summarize_launches.py gpurun_out/r2a_launches_ref_gpu.csv 0.5 > gpurun_out/r2a_launches_ref_gpu_summary.txt; head -14 gpurun_out/r2a_launches_ref_gpu_summary.txt
