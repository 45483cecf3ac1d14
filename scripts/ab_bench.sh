#!/bin/bash
# In-bench A/B of kernel-library builds (runs on the GPU box): scripts/ab_bench.sh <bench args> -- <lib or "main"> ...
# prints frames/s, attention / GEMM TFLOP/s and the SM clock of each build under the same power cap.
args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
for L in "$@"; do
  if [ "$L" == "main" ]; then unset DFOT_B200_LIB; else export DFOT_B200_LIB=$PWD/diffusion-forcing-transformer_b200/variants/lib_$L.so; fi
  python bench.py "${args[@]}" --skip-cpu-baseline --skip-parity 2>/dev/null | tail -1 > gpurun_out/ab_$L.json
  python - "$L" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
ra = d.get("roofline_attention", {})
print(f"{sys.argv[1]:10s} {d['value']:8.3f} {d['unit']}  attn {ra.get('achieved', 0):6.0f} TF share {ra.get('share_of_step', 0):.3f} "
      f"{ {k: round(v['tflops']) for k, v in ra.get('by_shape', {}).items()} }  gemm {d['roofline']['achieved']:6.0f} TF share "
      f"{d['roofline'].get('gemm_share_of_step', 0):.3f}  sm {d['clocks']['sm_mhz']:.0f} MHz")
PY
done
