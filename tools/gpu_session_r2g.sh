#!/bin/bash
# round-2 final single-GPU session: full GPU suite, the full bench line (reference-on-GPU leg, self-check), the image-text and LiT
# lines, micro-timings, the ncu launch list of one chunk and the ncu --set full capture of the forward GEMMs of a layer
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2g_smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 5000 gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
timeout 600 python bench.py --config image_text --steps 1 --warmup 1 > gpurun_out/r2g_bench_image_text.json 2> gpurun_out/r2g_bench_image_text.err; tail -c 1500 gpurun_out/r2g_bench_image_text.json; tail -3 gpurun_out/r2g_bench_image_text.err
timeout 600 python bench.py --config lit --steps 1 --warmup 1 > gpurun_out/r2g_bench_lit.json 2> gpurun_out/r2g_bench_lit.err; tail -c 1500 gpurun_out/r2g_bench_lit.json; tail -3 gpurun_out/r2g_bench_lit.err
timeout 300 python tools/bench_kernels.py > gpurun_out/r2g_kernels.log 2>&1; grep "^('gemm" gpurun_out/r2g_kernels.log | cut -c1-260
timeout 120 python tools/bench_attn.py 2>&1 | grep "bert\|vit" | cut -c1-300 | tee gpurun_out/r2g_attn.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_chunk.csv python tools/profile_chunk.py 2 > gpurun_out/r2g_prof_chunk.log 2>&1
python tools/summarize_launches.py gpurun_out/r2g_launches_chunk.csv 0.5 > gpurun_out/r2g_launches_chunk_summary.txt; head -18 gpurun_out/r2g_launches_chunk_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:gemm_kernel -s 48 -c 4 -o gpurun_out/r02g_gemm python tools/profile_chunk.py 1 > gpurun_out/r2g_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
