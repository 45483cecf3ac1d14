#!/bin/bash
# short bench run(s) on the GPU box; args are passed to bench.py
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "rc=$?"
tail -n 5 gpurun_out/bench.err
cat gpurun_out/bench.log
