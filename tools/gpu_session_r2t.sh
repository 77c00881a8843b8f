#!/bin/bash
# session T: schedule variants of attn_bwd4 (late dq_full wait, pipelined X loads, split drain): parity vs the product library + timing
mkdir -p gpurun_out
for v in "" latewait xpipe drainsplit latewait_xpipe all3; do
  CX_TRACE_VARIANT=$v timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2t_trace_${v:-plain}.log 2>&1
  echo "== ${v:-plain} rc=$?"
  grep -h "^variant\|^block_total\|worker0_tile2\|mma_tile2\|Error\|assert" gpurun_out/r2t_trace_${v:-plain}.log | cut -c1-300
done
