#!/bin/bash
set -u
export PYTHONUNBUFFERED=1
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_uvit_kernels.py -q -x --timeout 120 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 2>&1 | tail -2
show() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'frames/s', round(l['value'],2), 'NFE/s', round(l['nfe_per_sec'],1), 'ms/step', round(l['ms_per_step'],1), 'clk', l['clocks']['sm_mhz'])"; }
python bench.py --workload dmlab --frames 16 --batch 1 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T16 B1"
python bench.py --workload dmlab --frames 16 --batch 4 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T16 B4"
python bench.py --workload dmlab --frames 36 --batch 1 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T36 B1"
python bench.py --workload dmlab --frames 36 --batch 4 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T36 B4"
python bench.py --workload dmlab --frames 144 --batch 1 --steps 3 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "dmlab T144 B1"
python bench.py --workload k600 --batch 1 --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "k600 B1"
python bench.py --batch 1 --steps 2 --warmup 3 --skip-cpu-baseline 2>/dev/null | show "re10k B1"
