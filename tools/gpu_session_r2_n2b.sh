#!/bin/bash
# 2-GPU session B: the whole GPU suite (world-size-2 NCCL tests included), then the N=2 bench with the single all-reduce and with
# the bucketed reducer
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2_n2b_tests.log
timeout 100 python tools/bench_attn.py 2>&1 | grep "8 3" | cut -c1-200 | tee gpurun_out/r2_n2b_attn.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 2 --warmup 2 --no-selfcheck > gpurun_out/r2_n2b_bench.json 2> gpurun_out/r2_n2b_bench.err; tail -c 3000 gpurun_out/r2_n2b_bench.json; tail -3 gpurun_out/r2_n2b_bench.err
CX_OVERLAP_GRAD_REDUCE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 2 --steps 2 --warmup 2 --no-selfcheck --no-gpu-baseline > gpurun_out/r2_n2b_bench_overlap.json 2> gpurun_out/r2_n2b_bench_overlap.err; tail -c 1500 gpurun_out/r2_n2b_bench_overlap.json; tail -3 gpurun_out/r2_n2b_bench_overlap.err
