#!/bin/bash
# Profiling-only build of the library with the fast kernel's ablation switches (B200REC_ABL=<bits>, bpr_step.cu).
# Output: gpurun_build/libb200rec_abl.so (not the product; load with B200REC_LIB=...).
set -e
cd "$(dirname "$0")/../recsys_pytorch_b200/csrc"
mkdir -p ../../gpurun_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DB200REC_ABLATE -DB200REC_TC_DIAG -shared \
  capi.cu bpr_step.cu p2p.cu pointwise_step.cu score_exact.cu score_tc.cu metrics.cu spmm.cu ngcf.cu -o ../../gpurun_build/libb200rec_abl.so
echo built gpurun_build/libb200rec_abl.so
