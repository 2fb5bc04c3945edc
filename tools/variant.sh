#!/bin/bash
# Build a tuning variant of the library next to the product one: aac.js_b200/libaacfb_<name>.so
# (git-ignored, travels to the GPU box; select it with AACFB_LIB or tools/ab.sh).
# usage: tools/variant.sh <name> "<extra nvcc flags, e.g. -DAACFB_CS_ROT=0>"
set -e
cd "$(dirname "$0")/../aac.js_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off $2 \
     -shared -o ../libaacfb_$1.so aacfb_api.cu aacfb_kernels.cu aacfb_tables.cpp
