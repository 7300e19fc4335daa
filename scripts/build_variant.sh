#!/bin/bash
# builds a variant of the library with extra -D flags: scripts/build_variant.sh OUT.so -DFB_G3_DEPTH=5 ...
OUT=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared "$@" -o $OUT fitburst_b200/csrc/fb_api.cu
