#!/usr/bin/env python
"""Table of a `scripts/gpu.sh dmlab` sweep (one bench.py JSON line per configuration)."""
import json
import sys

print(f"{'frames':>6} {'tokens':>6} {'batch':>5} {'nfe':>3} {'frames/s':>10} {'NFE/s':>9} {'e2e fr/s':>10} {'ms/step':>9} "
      f"{'TFLOP/s':>8} {'gemm TF':>8} {'attn TF':>8}")
for line in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dmlab_sweep.jsonl"):
    try:
        l = json.loads(line)
    except Exception:
        continue
    w = l["config"]["workload"]
    T = int(w.split(" frames = ")[0].split()[-1])
    nfe = 2 if "nfe=2" in w else 1
    ra = l.get("roofline_attention") or {}
    print(f"{T:6d} {T * 16:6d} {l['config']['global_batch']:5d} {nfe:3d} {l['value']:10.1f} {l['nfe_per_sec']:9.1f} "
          f"{l['e2e']['value']:10.1f} {l['ms_per_step']:9.1f} {l['model_tflops']:8.1f} {l['roofline']['achieved']:8.1f} "
          f"{ra.get('achieved', 0):8.1f}")
