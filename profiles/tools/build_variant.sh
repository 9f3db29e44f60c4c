#!/bin/bash
# Build a variant of libmsort.so with extra nvcc flags (A/B experiments on one GPU box; selected with MSORT_LIB=...):
#   profiles/tools/build_variant.sh <name> [-DMACRO=value ...]   ->  marl-sortingenv_b200/csrc/variants/libmsort_<name>.so
set -e
cd "$(dirname "$0")/../../marl-sortingenv_b200/csrc"
name=$1; shift
mkdir -p variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v "$@" \
  -o variants/libmsort_$name.so msort_kernels.cu msort_policy.cu msort_ppo.cu msort_api.cu > variants/build_$name.log 2>&1 || { tail -20 variants/build_$name.log; exit 1; }
echo "built variants/libmsort_$name.so"
