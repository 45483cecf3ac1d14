#!/bin/bash
# The one GPU-side runner (everything here runs on the B200 box under gpurun; logs land in gpurun_out/).
#   scripts/gpu.sh round [nobench]          every -m gpu test file in its own process, smoke, the benches
#   scripts/gpu.sh tests <pytest args...>   python -m pytest -m gpu with the given selection
#   scripts/gpu.sh bench <bench.py args>    one bench.py line -> gpurun_out/bench_<tag>.json  (TAG=name to label it)
#   scripts/gpu.sh micro <attn|norm|sampler|uvit|gemm|...>   scripts/bench_kernels.py sections
#   scripts/gpu.sh prof <name> [name...]    ncu --set full capture of one kernel (names: see prof_cmd below)
#   scripts/gpu.sh launches [batch]         ncu launch list of two eager RE10K forwards
#   scripts/gpu.sh dmlab                    BASELINE config[4] sweep (T x batch)
#   scripts/gpu.sh ab <libA.so> <libB.so> <command...>   same command with two builds of the kernel library
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n "${TAILN:-3}" "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
  return $rc
}
prof_cmd() { # name -> kernel regex + command
  case $1 in
    attn64)         RX=attention; CMD="python scripts/bench_attn_one.py 8 9 64 8192 3 12.2" ;;
    attn64_maxpath) RX=attention; CMD="python scripts/bench_attn_one.py 8 9 64 8192 3" ;;
    attn72)         RX=attention; CMD="python scripts/bench_attn_one.py 8 16 72 1280 3" ;;
    attn128)        RX=attention; CMD="python scripts/bench_attn_one.py 8 9 128 2048 3 17.2" ;;
    conv)           RX="gemm2?_bf16"; CMD="python scripts/bench_one.py conv 3" ;;
    gemm)           RX="gemm2?_bf16"; CMD="python scripts/bench_one.py gemm 3" ;;
    gemm_l2)        RX="gemm2?_bf16"; CMD="python scripts/bench_one.py gemm_l2 3" ;;
    gn_silu)        RX=gn_silu; CMD="python scripts/bench_one.py gn_silu 3" ;;
    gn_stats)       RX=gn_stats; CMD="python scripts/bench_one.py gn_stats 3" ;;
    sampler)        RX=sampler; CMD="python scripts/bench_one.py sampler 3" ;;
    rmsnorm)        RX=rmsnorm_film; CMD="python scripts/bench_one.py rmsnorm 3" ;;
    rmsnorm1152)    RX=rmsnorm_film; CMD="python scripts/bench_one.py rmsnorm1152 3" ;;
    qknorm)         RX=qk_norm_rope; CMD="python scripts/bench_one.py qknorm 3" ;;
    adaln)          RX=adaln; CMD="python scripts/bench_one.py adaln 3" ;;
    patch_mix)      RX=patch_mix; CMD="python scripts/bench_one.py patch_mix 3" ;;
    patch_expand)   RX=patch_expand; CMD="python scripts/bench_one.py patch_expand 3" ;;
    pose_rays)      RX=pose_ray; CMD="python scripts/bench_pose_rays.py" ;;
    *) echo "unknown profile target $1"; return 1 ;;
  esac
}
task=${1:-round}; shift || true
case $task in
  round)
    : > gpurun_out/summary.txt
    nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
    for f in tests/test_gpu_*.py; do
      n=$(basename "$f" .py)
      run "$n" 1200 python -m pytest "$f" -q --timeout 900
    done
    run smoke 600 python __graft_entry__.py smoke
    if [ "${1:-}" != "nobench" ]; then
      TAILN=1 run bench_ref 600 python bench.py --impl reference --steps 1 --warmup 0
      TAILN=1 run bench_re10k 900 python bench.py --steps 3 --warmup 3
      TAILN=1 run bench_k600 900 python bench.py --workload k600 --steps 3 --warmup 3
    fi ;;
  tests)
    TAILN=15 run tests 2400 python -m pytest -m gpu -q --timeout 900 "$@" ;;
  bench)
    tag=${TAG:-$(echo "$*" | tr -c 'a-zA-Z0-9\n' '_' | cut -c1-60)}
    timeout 1500 python bench.py "$@" > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err"
    echo "rc=$?"; tail -c 3000 "gpurun_out/bench_${tag}.json"; tail -n 3 "gpurun_out/bench_${tag}.err" ;;
  micro)
    for sec in "$@"; do TAILN=40 run "micro_$sec" 600 python scripts/bench_kernels.py "$sec"; done ;;
  prof)
    for name in "$@"; do
      prof_cmd "$name" || continue
      $CMD > "gpurun_out/plain_$name.log" 2>&1 && \
      ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 2 -c 1 -o "gpurun_out/prof_$name" $CMD \
        > "gpurun_out/ncu_$name.log" 2>&1
      echo "$name rc=$? $(tail -1 gpurun_out/plain_$name.log)"
    done ;;
  launches)
    B=${1:-4}
    python scripts/profile_forward.py "$B" 2 > gpurun_out/fwd_plain.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_re10k.csv \
      python scripts/profile_forward.py "$B" 2 > gpurun_out/ncu_launch.log 2>&1
    echo "rc=$?"; cat gpurun_out/fwd_plain.log; tail -2 gpurun_out/ncu_launch.log ;;
  dmlab)
    out=gpurun_out/dmlab_sweep.jsonl; : > $out
    for T in 16 36 72 144; do for B in 1 4 16 64; do
      timeout 600 python bench.py --workload dmlab --frames $T --batch $B --steps 2 --warmup 3 --skip-cpu-baseline \
        2> gpurun_out/dmlab_err.log | tail -1 >> $out
      echo "T=$T B=$B rc=$?"
    done; done
    python scripts/dmlab_table.py $out ;;
  ab)
    A=$1; Bq=$2; shift 2
    for L in "$A" "$Bq"; do echo "--- $L"; DFOT_B200_LIB="$L" timeout 900 "$@" 2>&1 | tail -n "${TAILN:-6}"; done ;;
  *) echo "unknown task $task"; exit 2 ;;
esac
