#!/bin/bash
# usage: build_variant.sh NAME [-Dflags...]  -> scratch/variants/libcspm_NAME.so
set -e
cd "$(dirname "$0")/../crossscalepatchmatch_b200/csrc"
name=$1; shift
mkdir -p ../../scratch/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC,-ffp-contract=off,-O2 -Xptxas -v "$@" -shared -o ../../scratch/variants/libcspm_$name.so cspm_prep.cu cspm_pm.cu 2> ../../scratch/variants/$name.log
grep -A2 "k_refineILi0ELb0\|k_spatial_persistILi0ELi35" ../../scratch/variants/$name.log | grep "registers\|spill" | tr '\n' ' '; echo
