#!/bin/bash
# usage: profiles/build_f32_variant.sh NAME "-DFLAG=.. ..."  -> mtf_b200/csrc/_variants/libNAME.so
# only lk_ssd_f32.cu is recompiled (FCLK + Homography only); the other objects come from the main build (make -C mtf_b200/csrc)
set -e
cd "$(dirname "$0")/../mtf_b200/csrc"; mkdir -p _variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -DMTFB_ONLY_FCLK_HOM $2"
nvcc $FLAGS -Xptxas -v -c lk_ssd_f32.cu -o _variants/$1_f32.o 2> _variants/$1.lk_ssd_f32.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _variants/lib$1.so _variants/$1_f32.o _obj/lk_ssd.o _obj/lk_ncc.o _obj/lk_mi.o _obj/lk_mi_aff.o _obj/pf_kernels.o _obj/preproc.o _obj/mtfb_api.o
rm -f _variants/$1_f32.o
grep -A2 "ssd_update_f32_kernelILi0ELi1E" _variants/$1.lk_ssd_f32.ptxas.log | grep -E "Used|spill|Compiling" | sed 's/ptxas info    : //; s/Compiling entry function//; s/for .sm_100a.//'
