#!/bin/bash
# session H: phase trace of the attention backward (where a CTA's ~5700 clk per query tile go)
mkdir -p gpurun_out
timeout 300 python tools/trace_attn_bwd.py > gpurun_out/r2h_trace.log 2>&1
echo "trace rc=$?"
tail -60 gpurun_out/r2h_trace.log
