for cfg in "36 1" "16 4" "72 1"; do set -- $cfg; for mr in 512 1280; do
  echo "--- MAX_ROWS=$mr T=$1 B=$2"
  DFOT_DIT_SPLITK_MAX_ROWS=$mr timeout 600 python bench.py --workload dmlab --frames $1 --batch $2 --steps 3 --warmup 3 --skip-cpu-baseline --skip-parity 2>/dev/null | tail -1 | python -c "import json,sys; l=json.loads(sys.stdin.read()); print(round(l['value'],1), 'frames/s', round(l['ms_per_step'],2), 'ms/step')"
done; done
