#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/bench_kernels.py all > gpurun_out/kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/kernels.log
python scripts/bench_kernels.py attn --iters 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 5 -c 1 -o gpurun_out/prof_attn2 python scripts/bench_kernels.py attn --iters 2 > gpurun_out/ncu_attn2.log 2>&1
echo "ncu rc=$?"
