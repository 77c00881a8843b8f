#!/bin/bash
# session S: GradCache pass 2 seeded with the cached gradient directly (no torch.dot): the GradCache / training-step tests + smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_infonce.py tests/test_gpu_vit.py tests/test_gpu_poolers.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2s_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2s_smoke.log
