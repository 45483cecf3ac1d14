#!/bin/bash
# Programmatic dependent launch A/B: all GPU tests with PDL on, then both bench workloads with DFOT_PDL=1 and 0.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-3} "gpurun_out/$name.log"; }
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 120
run t_uvit_kernels 600 python -m pytest tests/test_gpu_uvit_kernels.py -q -x --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600
for pdl in 1 0; do
  for wl in re10k k600; do
    echo "=== bench $wl DFOT_PDL=$pdl"
    DFOT_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 2 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_${wl}_pdl$pdl.log 2>&1
    echo "rc=$?"; python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_${wl}_pdl$pdl.log").read().strip().splitlines()[-1])
    print("value", round(l["value"],3), "nfe/s", round(l["nfe_per_sec"],2), "clk", l["clocks"]["sm_mhz"], "gemm", round(l["roofline"]["achieved"],1), "attn", round(l["roofline_attention"]["achieved"],1))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_${wl}_pdl$pdl.log").read()[-1500:])
PY
  done
done
