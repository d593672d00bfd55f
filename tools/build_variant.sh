#!/bin/bash
# A/B builds of the native library: tools/build_variant.sh NAME [-DFLAG=VALUE ...]  ->  build_variants/libNAME.so
# (select at run time with OCL_SC_LIB=build_variants/libNAME.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
name=$1; shift
src=ocelot_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -shared -o build_variants/lib$name.so "$@" \
  $src/sc_kernels.cu $src/sc_fft.cu $src/sc_beam.cu $src/sc_lsc.cu $src/sc_abi.cu -lcufft \
  -Xlinker -rpath=/usr/local/cuda/lib64
echo built build_variants/lib$name.so
