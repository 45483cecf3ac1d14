#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-3} "gpurun_out/$name.log"; }
run t_attn_a 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention" --timeout 120
DFOT_B200_LIB=$PWD/diffusion-forcing-transformer_b200/build/lib_nopack.so run t_attn_b 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention" --timeout 120
