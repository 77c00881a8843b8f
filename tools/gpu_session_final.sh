#!/bin/bash
# round-end style measurement: bench.py (N=1), the ncu launch list of one GradCache chunk, ncu --set full of the top kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_chunk.csv 0.5 > gpurun_out/launches_chunk_summary.txt; head -12 gpurun_out/launches_chunk_summary.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"attn_(fwd3|bwd2)" -c 2 -o gpurun_out/r01_attn_final python tools/bench_attn.py 6,2 > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -s 48 -c 4 -o gpurun_out/r01_gemm_final python tools/profile_chunk.py 1 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out | tail -12
