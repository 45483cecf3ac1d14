#!/bin/bash
# K600 (DiT-XL) small-batch A/B: latency mode off / on (bench.py --latency-mode), batch 1 (1280 token rows) and 2
for B in 1 2; do for sk in 0 1; do
  echo "--- latency_mode=$sk K600 B=$B"; flag=""; [ $sk = 1 ] && flag="--latency-mode"
  timeout 600 python bench.py --workload k600 --batch $B --steps 3 --warmup 3 --skip-cpu-baseline --skip-parity $flag 2>/dev/null | tail -1 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'frames/s', round(l['ms_per_step'],2), 'ms/step', round(l['e2e']['value'],1), 'e2e')"
done; done
