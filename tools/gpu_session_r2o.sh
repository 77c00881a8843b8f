#!/bin/bash
# session O: attn_bwd4 with dQ issued one slot late: parity, phase trace, timing against bwd3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_ops.py -q -x -k "attention" > gpurun_out/r2o_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2o_tests.log
timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2o_trace_bwd4.log 2>&1
echo "== trace rc=$?"; cut -c1-300 gpurun_out/r2o_trace_bwd4.log | tail -30
timeout 300 python tools/bench_attn.py > gpurun_out/r2o_bench_attn_bwd4.log 2>&1; echo "bench5 rc=$?"
CX_ATTN_BWD3=1 timeout 300 python tools/bench_attn.py > gpurun_out/r2o_bench_attn_bwd3.log 2>&1; echo "bench3 rc=$?"
grep -h "bert_64x512 8\|vit_256x197 8" gpurun_out/r2o_bench_attn_bwd4.log gpurun_out/r2o_bench_attn_bwd3.log | cut -c1-200
