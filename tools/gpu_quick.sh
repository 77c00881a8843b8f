cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_encoder_ops.py tests/test_gpu_encoder.py -q -m gpu 2>&1 | tail -3
timeout 60 ./tools/ubench/ubench_tmem 2>&1 | tee gpurun_out/ubench4.log | grep UTCHMMA
