#!/bin/bash
# 2-GPU session C (final state): world-size-2 NCCL tests + the tests touched last, N=2 bench, LiT line on one GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_poolers.py tests/test_gpu_vit.py tests/test_gpu_infonce.py -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r2_n2c_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 2 --warmup 2 --no-selfcheck --no-gpu-baseline > gpurun_out/r2_n2c_bench.json 2> gpurun_out/r2_n2c_bench.err; tail -c 1600 gpurun_out/r2_n2c_bench.json; tail -3 gpurun_out/r2_n2c_bench.err
timeout 600 python bench.py --config lit --steps 1 --warmup 1 > gpurun_out/r2_n2c_bench_lit.json 2> gpurun_out/r2_n2c_bench_lit.err; tail -c 1200 gpurun_out/r2_n2c_bench_lit.json; tail -3 gpurun_out/r2_n2c_bench_lit.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference > gpurun_out/r2_n2c_ref_arm.json 2> gpurun_out/r2_n2c_ref_arm.err; tail -c 700 gpurun_out/r2_n2c_ref_arm.json; tail -2 gpurun_out/r2_n2c_ref_arm.err
