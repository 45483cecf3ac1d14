#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
V=diffusion-forcing-transformer_b200/build/variants
for v in p12 p19 p25 p31 p37 p44 p50 p25 p50; do
  for cs in 0; do
    if [ $v = base ]; then lib=""; else lib="$PWD/$V/libdfot_$v.so"; fi
    echo "--- $v"
    DFOT_B200_LIB=$lib DFOT_ATTN_COLSPLIT=$cs python scripts/bench_kernels.py attn 2>&1 | grep -A1 "R=8 heads=9 d=64 N=8192" | tail -1
  done
done
