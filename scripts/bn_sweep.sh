for bn in 0 128 192 256; do
  echo "##### DFOT_GEMM_BN=$bn"
  if [ $bn = 0 ]; then python scripts/bench_kernels.py gemm; python scripts/bench_kernels.py uvit_gemm;
  else DFOT_GEMM_BN=$bn python scripts/bench_kernels.py gemm; DFOT_GEMM_BN=$bn python scripts/bench_kernels.py uvit_gemm; fi
done > gpurun_out/bn_sweep.txt 2>&1
