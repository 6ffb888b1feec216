#!/bin/bash
# usage: tools/build_variant.sh NAME "<extra nvcc flags>"  ->  tools/_build/libevrp_NAME.so (tuning variants; run with EVRP_B200_LIB=...)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../drl-energy-optimal-routing-for-electric-vehicles_b200/csrc"
mkdir -p ../../tools/_build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
  -shared -o ../../tools/_build/libevrp_$NAME.so evrp_step.cu evrp_encoder.cu evrp_encoder_fused.cu evrp_decode.cu evrp_decode_tc.cu evrp_attn_tc.cu evrp_tc_gemm.cu evrp_loss.cu evrp_logp.cu evrp_attn_train.cu evrp_wgrad.cu evrp_train.cu -lcudart
echo built tools/_build/libevrp_$NAME.so
