#!/bin/bash
# Build an experiment variant of the product library: tools/dev/build_variant.sh NAME [-DFLAG=1 ...]
# -> cutter-vad_b200/variants/libcvad_NAME.so (git-ignored); select it with CVAD_B200_LIB=<path>.
set -e
here=$(cd "$(dirname "$0")/../.." && pwd)
name=$1; shift
mkdir -p "$here/cutter-vad_b200/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
     -o "$here/cutter-vad_b200/variants/libcvad_$name.so" "$here/cutter-vad_b200/csrc/cvad_capi.cu"
echo "$here/cutter-vad_b200/variants/libcvad_$name.so"
