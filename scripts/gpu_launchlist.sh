#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/profile_forward.py 4 2 > gpurun_out/fwd_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_re10k.csv python scripts/profile_forward.py 4 2 > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"; cat gpurun_out/fwd_plain.log; tail -2 gpurun_out/ncu_launch.log
