#!/bin/bash
# PDL on/off at the launch-bound end of the DMLab sweep and on the default workload.
set -u
export PYTHONUNBUFFERED=1
show() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'frames/s', round(l['value'],2), 'NFE/s', round(l['nfe_per_sec'],1), 'ms/step', round(l['ms_per_step'],1), 'clk', l['clocks']['sm_mhz'])"; }
for pdl in 0 1 0 1; do
  DFOT_PDL=$pdl python bench.py --workload dmlab --frames 16 --batch 1 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T16 B1 pdl=$pdl"
  DFOT_PDL=$pdl python bench.py --workload dmlab --frames 36 --batch 4 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T36 B4 pdl=$pdl"
done
for pdl in 0 1; do
  DFOT_PDL=$pdl python bench.py --workload k600 --batch 1 --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "k600 B1 pdl=$pdl"
  DFOT_PDL=$pdl python bench.py --batch 1 --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "re10k B1 pdl=$pdl"
done
