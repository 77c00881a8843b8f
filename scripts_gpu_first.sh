#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
# forward GEMMs of layer 0 in the grad-mode forward: skip the 48 no-grad forward GEMMs first
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 48 -c 4 -o gpurun_out/prof_gemm python tools/profile_chunk.py 1 > gpurun_out/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 12 -c 1 -o gpurun_out/prof_attn_fwd python tools/profile_chunk.py 1 >> gpurun_out/prof_gemm.log 2>&1
tail -3 gpurun_out/prof_gemm.log
