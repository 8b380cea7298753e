#!/bin/bash
# usage: profiles/build_variant.sh NAME "-DFLAG=.. ..."  -> mtf_b200/csrc/_variants/libNAME.so (FCLK+Homography only)
set -e
cd "$(dirname "$0")/../mtf_b200/csrc"; mkdir -p _variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -DMTFB_ONLY_FCLK_HOM $2"
nvcc $FLAGS -Xptxas -v -c lk_kernels.cu -o _variants/$1_k.o 2> _variants/$1.ptxas.log
nvcc $FLAGS -c mtfb_api.cu -o _variants/$1_a.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _variants/lib$1.so _variants/$1_k.o _variants/$1_a.o
rm -f _variants/$1_k.o _variants/$1_a.o
grep -A2 "lk_update_kernelILi0ELi0ELi1ELi64E" _variants/$1.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    : //'
