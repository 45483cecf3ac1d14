#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/bench_attn_one.py 8 9 64 8192 3 40 > gpurun_out/attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -o gpurun_out/prof_attn_nomax python scripts/bench_attn_one.py 8 9 64 8192 3 40 > gpurun_out/ncu_attn_nomax.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_nomax.log
