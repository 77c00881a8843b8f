#!/bin/bash
# session J: ablations of the attention backward (each variant removes one consumer of a shared resource; results are wrong on purpose)
mkdir -p gpurun_out
for v in "" nostat nodsstore noexp nodrain nodqmma nostat_noexp all_off; do
  CX_TRACE_VARIANT=$v timeout 200 python tools/trace_attn_bwd.py > gpurun_out/r2j_trace_${v:-plain}.log 2>&1
  echo "== ${v:-plain} rc=$?"
  grep -h "^variant\|^block_total\|worker0_tile1\|mma_tile1\|worker_epilogue" gpurun_out/r2j_trace_${v:-plain}.log | cut -c1-300
done
