#!/bin/bash
# Build a library variant into scratch_libs/<name>.so with extra nvcc flags: scripts/build_variant.sh <name> [-DFLAG ...]
NAME=$1; shift
mkdir -p scratch_libs
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o scratch_libs/${NAME}.so windgym_b200/csrc/api.cu windgym_b200/csrc/flow.cu windgym_b200/csrc/env.cu windgym_b200/csrc/field.cu
