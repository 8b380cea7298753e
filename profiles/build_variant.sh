#!/bin/bash
# usage: profiles/build_variant.sh NAME "-DFLAG=.. ..."  -> mtf_b200/csrc/_variants/libNAME.so (FCLK+Homography only)
set -e
cd "$(dirname "$0")/../mtf_b200/csrc"; mkdir -p _variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -DMTFB_ONLY_FCLK_HOM $2"
for f in lk_ssd lk_ssd_f32 lk_ncc lk_mi lk_mi_aff pf_kernels preproc mtfb_api; do nvcc $FLAGS -Xptxas -v -c $f.cu -o _variants/$1_$f.o 2> _variants/$1.$f.ptxas.log & done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _variants/lib$1.so _variants/$1_*.o
rm -f _variants/$1_*.o
grep -A2 "ssd_update_kernelILi0ELi1ELi32E" _variants/$1.lk_ssd.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    : //'
grep -A2 "ssd_update_f32_kernelILi0ELi1ELi64E" _variants/$1.lk_ssd_f32.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    : //'
