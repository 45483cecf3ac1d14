#!/bin/bash
# DMLab small-batch A/B: latency mode off / on (bench.py --latency-mode), T = 16 at batch 1 / 2 / 4 and T = 36 at batch 1
for cfg in "16 1" "16 2" "36 1" "16 4"; do set -- $cfg; for sk in 0 1; do
  echo "--- latency_mode=$sk T=$1 B=$2"; flag=""; [ $sk = 1 ] && flag="--latency-mode"
  timeout 600 python bench.py --workload dmlab --frames $1 --batch $2 --steps 3 --warmup 3 --skip-cpu-baseline --skip-parity $flag 2>/dev/null | tail -1 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'frames/s', round(l['ms_per_step'],2), 'ms/step', round(l['e2e']['value'],1), 'e2e', l['gpu_launches'], 'launches')"
done; done
