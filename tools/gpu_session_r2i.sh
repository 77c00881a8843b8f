#!/bin/bash
# session I: four-threads-per-row attention backward (attn_bwd4): parity tests, phase trace, timing against bwd3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_ops.py -q -x -k "attention" > gpurun_out/r2i_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2i_tests.log
timeout 300 python tools/trace_attn_bwd.py > gpurun_out/r2i_trace_bwd4.log 2>&1
echo "trace rc=$?"; head -12 gpurun_out/r2i_trace_bwd4.log; grep mma_ gpurun_out/r2i_trace_bwd4.log
timeout 300 python tools/bench_attn.py > gpurun_out/r2i_bench_attn_bwd4.log 2>&1; echo "bench4 rc=$?"
cp gpurun_out/bench_attn_fwd4_bwd3.json gpurun_out/r2i_bench_attn_bwd4.json
CX_ATTN_BWD3=1 timeout 300 python tools/bench_attn.py > gpurun_out/r2i_bench_attn_bwd3.log 2>&1; echo "bench3 rc=$?"
cp gpurun_out/bench_attn_fwd4_bwd3.json gpurun_out/r2i_bench_attn_bwd3.json
grep -h "bert_64x512\|vit_256" gpurun_out/r2i_bench_attn_bwd4.log gpurun_out/r2i_bench_attn_bwd3.log | cut -c1-330
