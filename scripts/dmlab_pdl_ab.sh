#!/bin/bash
# DMLab small-batch A/B: programmatic dependent launch off / on (DFOT_PDL), batch 1 and 4 at T = 16
for pdl in 0 1; do for B in 1 4; do
  echo "--- DFOT_PDL=$pdl T=16 B=$B"
  DFOT_PDL=$pdl timeout 600 python bench.py --workload dmlab --frames 16 --batch $B --steps 3 --warmup 3 --skip-cpu-baseline --skip-parity 2>/dev/null | tail -1 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(l['value'], 'frames/s', l['ms_per_step'], 'ms/step', l['e2e']['value'], 'e2e')"
done; done
