#!/bin/bash
# Runs on the GPU box via gpurun: kernel unit tests (one process per kernel family so a trapped kernel cannot
# poison the others), parity tests, smoke, a short bench.  Logs land in gpurun_out/.
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
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_glue 300 python -m pytest tests/test_gpu_kernels.py -q -k "noise_features" --timeout 120
run t_sampler 300 python -m pytest tests/test_gpu_kernels.py -q -k "sampler_step" --timeout 120
run t_adaln 300 python -m pytest tests/test_gpu_kernels.py -q -k "adaln" --timeout 120
run t_attn 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention" --timeout 120
run t_gemm 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm_bias" --timeout 120
run t_gemm_epi 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm_activations or gemm_qkv" --timeout 120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout 300
run smoke 300 python __graft_entry__.py smoke
if [ "${1:-}" != "nobench" ]; then
  run bench 900 python bench.py --steps 2 --warmup 3
fi
cat gpurun_out/summary.txt
