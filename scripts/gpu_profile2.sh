#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --sampling-steps 2 --skip-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 10 -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 60 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
