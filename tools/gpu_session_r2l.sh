#!/bin/bash
# session L: cheap stamps (pointer read once): faithful phase trace of bwd4 and bwd3, and the ablations timed after a clock ramp
mkdir -p gpurun_out
for v in "" nodrain nostat nodsstore noexp nodqmma all_off; do
  CX_TRACE_VARIANT=$v timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2l_trace_${v:-plain}.log 2>&1
  echo "== ${v:-plain} rc=$?"
  grep -h "^variant\|^block_total\|worker0_tile\|mma_tile\|drain_tile1\|worker_epilogue\|acc_full\|drain_done\|setup" gpurun_out/r2l_trace_${v:-plain}.log | cut -c1-300
done
CX_ATTN_BWD3=1 timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2l_trace_bwd3.log 2>&1
echo "== bwd3 rc=$?"; grep -h "^variant\|^block_total\|worker0_tile\|mma_tile\|worker_epilogue\|acc_full\|drain_done\|setup" gpurun_out/r2l_trace_bwd3.log | cut -c1-300
