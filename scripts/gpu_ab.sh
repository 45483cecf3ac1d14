#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== default (no poly)"; python -m pytest tests/test_gpu_kernels.py -q -k "bounded" --timeout 120 2>&1 | tail -1
python scripts/bench_kernels.py attn 2>&1 | grep -A1 "R=8 heads=9 d=64 N=8192\|d=64 N=1280" | grep bounded
for v in a b c; do
  echo "== variant $v"
  DFOT_B200_LIB=$PWD/diffusion-forcing-transformer_b200/build/lib_polyb_$v.so python -m pytest tests/test_gpu_kernels.py -q -k "bounded" --timeout 120 2>&1 | tail -1
  DFOT_B200_LIB=$PWD/diffusion-forcing-transformer_b200/build/lib_polyb_$v.so python scripts/bench_kernels.py attn 2>&1 | grep -A1 "R=8 heads=9 d=64 N=8192\|d=64 N=1280" | grep bounded
done
