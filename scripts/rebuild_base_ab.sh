#!/bin/bash
# A/B of DFOT_DIT_REBUILD_BASE (K1 stores bf16 + row statistics, the gate epilogue rebuilds the fp32 residual base in place) on a DiT workload
W=${1:-k600}; shift || true
for v in 0 1; do
  echo "--- $W DFOT_DIT_REBUILD_BASE=$v $*"
  DFOT_DIT_REBUILD_BASE=$v timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --skip-cpu-baseline --skip-parity "$@" 2>/dev/null | tail -1 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(round(l['value'],2), 'frames/s', round(l['ms_per_step'],2), 'ms/step', round(l['e2e']['value'],2), 'e2e', l['clocks']['sm_mhz'], 'MHz', 'gemm', round(l['roofline']['frac'],3), round(l['roofline']['gemm_share_of_step'],3))"
done
