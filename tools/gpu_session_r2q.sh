#!/bin/bash
# round-2 closing single-GPU session (after the attention-backward re-schedule): full GPU suite, smoke, the full bench line,
# attention timings, the launch list of one chunk and an ncu --set full capture of the backward attention kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2q_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2q_smoke.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 4000 gpurun_out/r2q_bench.json; tail -3 gpurun_out/r2q_bench.err
timeout 120 python tools/bench_attn.py 2>&1 | grep "bert\|vit" | cut -c1-300 | tee gpurun_out/r2q_attn.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2q_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2q_launches_chunk.csv 0.5 > gpurun_out/r2q_launches_chunk_summary.txt; head -18 gpurun_out/r2q_launches_chunk_summary.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_bwd4 -s 2 -c 1 -o gpurun_out/r02q_attn_bwd4 python tools/bench_attn.py > gpurun_out/r2q_ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
