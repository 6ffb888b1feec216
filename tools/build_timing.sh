#!/bin/bash
# instrumented debug build of the library (phase cycle counters in the rollout kernel); not the product .so
set -e
cd "$(dirname "$0")/../drl-energy-optimal-routing-for-electric-vehicles_b200/csrc"
mkdir -p ../../tools/_build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DEVRP_PHASE_TIMING \
  -shared -o ../../tools/_build/libevrp_b200_timing.so evrp_step.cu evrp_encoder.cu evrp_encoder_fused.cu evrp_decode.cu evrp_decode_tc.cu evrp_attn_tc.cu evrp_tc_gemm.cu evrp_loss.cu evrp_logp.cu evrp_attn_train.cu evrp_wgrad.cu evrp_train.cu -lcudart
echo built tools/_build/libevrp_b200_timing.so
