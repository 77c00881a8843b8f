#!/bin/bash
# round-2 session E: reducer plumbing reproduced on one GPU (2 gloo ranks), single-accumulation Matryoshka, activation retention,
# SwiGLU-backward prefetch; ncu --set full of the attention backward; bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multirank.py -q -m gpu -x -k "one_gpu" 2>&1 | tail -45 | tee gpurun_out/r2e_reducer_one_gpu.log
timeout 400 python -m pytest tests/test_gpu_infonce.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/r2e_infonce.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r2e_tests.log
timeout 300 python tools/bench_kernels.py > gpurun_out/r2e_kernels.log 2>&1; grep "swiglu_bwd\|infonce" gpurun_out/r2e_kernels.log | cut -c1-300
timeout 200 python tools/bench_matryoshka.py 2>&1 | tail -8 | tee gpurun_out/r2e_matryoshka.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd3 -s 2 -c 1 -o gpurun_out/r02e_attn_bwd python tools/bench_attn.py > gpurun_out/r2e_ncu_attn.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-baseline --no-selfcheck > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 1200 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
CX_RETAIN_ACTIVATIONS=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-gpu-baseline --no-selfcheck > gpurun_out/r2e_bench_noretain.json 2> gpurun_out/r2e_bench_noretain.err; tail -c 600 gpurun_out/r2e_bench_noretain.json | head -c 300
