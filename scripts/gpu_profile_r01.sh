#!/bin/bash
# Round-1 evidence: per-kernel ncu --set full captures of the dominant kernels at RE10K sizes (each after a plain run
# of the same command exited 0) and the launch list of one RE10K forward pair.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
prof() { # name, kernel regex, command...   (ONLY="conv gemm" restricts the captures)
  local name=$1 rx=$2; shift 2
  if [ -n "${ONLY:-}" ] && ! [[ " $ONLY " == *" $name "* ]]; then return; fi
  "$@" > "gpurun_out/plain_$name.log" 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 2 -c 1 -o "gpurun_out/prof_$name" "$@" > "gpurun_out/ncu_$name.log" 2>&1
  echo "$name rc=$? $(cat gpurun_out/plain_$name.log | tail -1)"
}
prof attn64 attention python scripts/bench_attn_one.py 8 9 64 8192 3 12.2
prof attn64_maxpath attention python scripts/bench_attn_one.py 8 9 64 8192 3
prof attn128 attention python scripts/bench_attn_one.py 8 9 128 2048 3 17.2
prof conv "gemm2?_bf16" python scripts/bench_one.py conv 3
prof gemm "gemm2?_bf16" python scripts/bench_one.py gemm 3
prof gemm_l2 "gemm2?_bf16" python scripts/bench_one.py gemm_l2 3
prof gn_silu gn_silu python scripts/bench_one.py gn_silu 3
prof sampler sampler python scripts/bench_one.py sampler 3
prof rmsnorm rmsnorm_film python scripts/bench_one.py rmsnorm 3
prof qknorm qk_norm_rope python scripts/bench_one.py qknorm 3
[ -n "${ONLY:-}" ] && exit 0
python scripts/profile_forward.py 4 2 > gpurun_out/fwd_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_re10k.csv python scripts/profile_forward.py 4 2 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
