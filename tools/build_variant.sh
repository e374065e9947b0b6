#!/bin/bash
# Build an A/B variant of libnerfsos.so: tools/build_variant.sh NAME [-DFLAG ...]  ->  nerf-sos_b200/lib/libnerfsos_NAME.so
# (select it at run time with NSOS_LIB=nerf-sos_b200/lib/libnerfsos_NAME.so)
set -e
cd "$(dirname "$0")/../nerf-sos_b200"
name=$1; shift
mkdir -p lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -o lib/libnerfsos_${name}.so \
  csrc/api.cu csrc/simt_gemm.cu csrc/simt_render.cu csrc/corr_loss.cu csrc/tc_render.cu csrc/tc_wgrad.cu csrc/optim.cu
echo built lib/libnerfsos_${name}.so
