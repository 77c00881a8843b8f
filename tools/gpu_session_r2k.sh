#!/bin/bash
# session K: attn_bwd4 with a 3-stage Q/dO ring: parity, trace (plain + two ablations), timing against bwd3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder_ops.py -q -x -k "attention" > gpurun_out/r2k_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
for v in "" nodrain nostat all_off; do
  CX_TRACE_VARIANT=$v timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2k_trace_${v:-plain}.log 2>&1
  echo "== ${v:-plain} rc=$?"
  grep -h "^variant\|^block_total\|worker0_tile\|mma_tile\|worker_epilogue\|acc_full\|drain_done" gpurun_out/r2k_trace_${v:-plain}.log | cut -c1-300
done
timeout 300 python tools/bench_attn.py > gpurun_out/r2k_bench_attn_bwd4.log 2>&1; echo "bench4 rc=$?"
CX_ATTN_BWD3=1 timeout 300 python tools/bench_attn.py > gpurun_out/r2k_bench_attn_bwd3.log 2>&1; echo "bench3 rc=$?"
grep -h "bert_64x512 8\|vit_256x197 8" gpurun_out/r2k_bench_attn_bwd4.log gpurun_out/r2k_bench_attn_bwd3.log | cut -c1-200
