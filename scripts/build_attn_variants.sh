#!/bin/bash
# Build variant libraries of the attention kernel (polynomial-exp2 share of the bounded-score path) for A/B runs:
#   build/variants/libdfot_<name>.so, selected at run time with DFOT_B200_LIB.
set -eu
cd "$(dirname "$0")/../diffusion-forcing-transformer_b200"
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OTHERS=$(ls build/*.o | grep -v "attention")
build() { # name first second
  nvcc $FLAGS -DDFOT_ATTN_POLYB_FIRST=$2 -DDFOT_ATTN_POLYB_SECOND=$3 -c csrc/attention_tcgen05.cu -o build/variants/attn_$1.o
  nvcc -shared -o build/variants/libdfot_$1.so build/variants/attn_$1.o $OTHERS -Xcompiler -fPIC -cudart static
}
build p12 0x00 0x11 &
build p19 0x00 0x15 &
build p25 0x00 0x55 &
build p31 0x00 0x57 &
wait
build p37 0x00 0x77 &
build p44 0x00 0x7F &
build p50 0x00 0xFF &
wait
ls -la build/variants/*.so
