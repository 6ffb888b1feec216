#!/bin/bash
# debug build of the library with the fused encoder's wait-time profile compiled in -> tools/_build/libevrp_prof.so
set -e
cd "$(dirname "$0")/../drl-energy-optimal-routing-for-electric-vehicles_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DEVRP_FENC_PROF $EXTRA \
  -shared -o ../../tools/_build/libevrp_prof.so evrp_step.cu evrp_encoder.cu evrp_encoder_fused.cu evrp_decode.cu evrp_decode_tc.cu evrp_attn_tc.cu evrp_tc_gemm.cu evrp_loss.cu evrp_logp.cu evrp_attn_train.cu evrp_wgrad.cu evrp_train.cu -lcudart
echo built tools/_build/libevrp_prof.so
