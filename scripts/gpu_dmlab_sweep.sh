#!/bin/bash
# BASELINE config[4]: DMLab-shaped long-context throughput sweep (DiT-B, T in {16,36,72,144} frames, batch 1-64) on one GPU.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
out=gpurun_out/dmlab_sweep.jsonl
: > $out
for T in 16 36 72 144; do
  for B in 1 4 16 64; do
    timeout 600 python bench.py --workload dmlab --frames $T --batch $B --steps 2 --warmup 3 --skip-cpu-baseline 2> gpurun_out/dmlab_err.log | tail -1 >> $out
    echo "T=$T B=$B rc=$?"
  done
done
timeout 600 python bench.py --workload dmlab --frames 36 --batch 16 --guidance 2.0 --steps 2 --warmup 3 2>> gpurun_out/dmlab_err.log | tail -1 >> $out
python - <<'PY'
import json
print(f"{'frames':>6} {'tokens':>6} {'batch':>5} {'nfe':>3} {'frames/s':>10} {'NFE/s':>9} {'e2e fr/s':>10} {'ms/step':>9} {'TFLOP/s':>8} {'gemm TF':>8} {'attn TF':>8}")
for line in open("gpurun_out/dmlab_sweep.jsonl"):
    try:
        l = json.loads(line)
    except Exception:
        continue
    w = l["config"]["workload"]
    T = int(w.split(" frames = ")[0].split()[-1])
    nfe = 2 if "nfe=2" in w else 1
    B = l["config"]["global_batch"]
    ra = l.get("roofline_attention") or {}
    print(f"{T:6d} {T*16:6d} {B:5d} {nfe:3d} {l['value']:10.1f} {l['nfe_per_sec']:9.1f} {l['e2e']['value']:10.1f} {l['ms_per_step']:9.1f} {l['model_tflops']:8.1f} {l['roofline']['achieved']:8.1f} {ra.get('achieved', 0):8.1f}")
PY
