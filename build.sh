#!/bin/bash
# Builds libtgp_b200.so in-tree for sm_100a (the .so is git-ignored but travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/tgp/pytorch_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     -o ../libtgp_b200.so tgp_b200.cu "$@"
