#!/bin/bash
# build_variant.sh NAME [extra nvcc flags...]: an alternative build of libdajin2_b200.so for kernel experiments
# (dajin2_b200/variants/lib<NAME>.so, selected at run time with DAJIN2_B200_LIB).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p dajin2_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -shared --cudart shared \
  -o dajin2_b200/variants/lib${name}.so dajin2_b200/csrc/cabi.cu dajin2_b200/csrc/encode.cu dajin2_b200/csrc/encode_v1.cu dajin2_b200/csrc/hist.cu \
  dajin2_b200/csrc/kmer.cu dajin2_b200/csrc/cluster.cu dajin2_b200/csrc/readstat.cu 2>&1 | grep -v "deprecated-gpu-targets" || true
ls -la dajin2_b200/variants/lib${name}.so
