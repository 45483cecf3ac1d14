#!/bin/bash
# Full GPU round: every -m gpu test file in its own process (a trapped kernel cannot poison the others), smoke, both bench
# workloads.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 3 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_kernels 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 120
run t_uvit_kernels 600 python -m pytest tests/test_gpu_uvit_kernels.py -q --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout 600
run t_vae_kernels 600 python -m pytest tests/test_gpu_vae_kernels.py -q --timeout 120
run t_vae_decode 600 python -m pytest tests/test_gpu_vae_decode.py -q --timeout 300
run smoke 600 python __graft_entry__.py smoke
if [ "${1:-}" != "nobench" ]; then
  run bench_re10k 900 python bench.py --steps 3 --warmup 3
  run bench_k600 900 python bench.py --workload k600 --steps 3 --warmup 3
  run bench_ref 600 python bench.py --impl reference --steps 1 --warmup 0
  run bench_vae 300 python scripts/bench_vae_decode.py --batch 8
fi
