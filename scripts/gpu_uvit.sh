#!/bin/bash
# U-ViT3DPose bring-up on the GPU box: new kernel tests (one process per family), U-ViT parity, regression of the GEMM.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "rc=$?"; tail -n ${TAILN:-12} "gpurun_out/$name.log"; }
run u_conv 300 python -m pytest tests/test_gpu_uvit_kernels.py -q -k "conv3x3 or gemm_resid" --timeout 120
run u_norm 300 python -m pytest tests/test_gpu_uvit_kernels.py -q -k "groupnorm or rmsnorm or qk_norm or pool" --timeout 120
run u_pose 300 python -m pytest tests/test_gpu_uvit_kernels.py -q -k "pose_ray" --timeout 120
run u_fwd 600 python -m pytest tests/test_gpu_parity.py -q -k "uvit" --timeout 300
run t_gemm 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" --timeout 120
