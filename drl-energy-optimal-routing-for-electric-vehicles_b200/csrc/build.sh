#!/bin/bash
# Build libevrp_b200.so in-tree for sm_100a (no other arch, no fallback).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libevrp_b200.so
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v \
  -shared -o $OUT evrp_step.cu evrp_encoder.cu evrp_encoder_fused.cu evrp_decode.cu evrp_decode_tc.cu evrp_attn_tc.cu evrp_tc_gemm.cu evrp_loss.cu evrp_logp.cu evrp_attn_train.cu evrp_wgrad.cu evrp_train.cu -lcudart 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log || true
echo "built $OUT"
