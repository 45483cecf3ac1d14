#!/bin/bash
# Build an A/B variant of the kernel library: the named .cu is recompiled with extra -D flags, everything else is re-used.
#   scripts/build_variant.sh <name> <source.cu> [-DFLAG ...]   ->  diffusion-forcing-transformer_b200/variants/lib_<name>.so
# Use with DFOT_B200_LIB=<that file> (scripts/gpu.sh ab ...).  variants/ is git-ignored and travels with the snapshot.
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)/diffusion-forcing-transformer_b200
python -c "import sys; sys.path.insert(0, '$root/..'); from dfot_b200.build import build; build()"
mkdir -p "$root/variants"
obj="$root/variants/${name}_${src}.o"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
  -c "$root/csrc/$src" -o "$obj"
objs=""
for f in "$root"/build/*.cu.o; do
  if [ "$(basename "$f")" == "$src.o" ]; then objs="$objs $obj"; else objs="$objs $f"; fi
done
nvcc -shared -o "$root/variants/lib_${name}.so" $objs -Xcompiler -fPIC -cudart static
echo "$root/variants/lib_${name}.so"
