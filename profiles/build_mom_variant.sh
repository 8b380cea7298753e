#!/bin/bash
# usage: profiles/build_mom_variant.sh NAME "-DFLAG=.. ..."  -> scratch/variants/libNAME.so
# lk_ssd_mom.cu and mtfb_api.cu are recompiled with the flags; the other objects come from the main build (make -C mtf_b200/csrc)
set -e
cd "$(dirname "$0")/../mtf_b200/csrc"; V=../../scratch/variants; mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off $2"
nvcc $FLAGS -Xptxas -v -c lk_ssd_mom.cu -o $V/$1_mom.o 2> $V/$1.lk_ssd_mom.ptxas.log &
nvcc $FLAGS -c mtfb_api.cu -o $V/$1_api.o 2> $V/$1.api.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib$1.so $V/$1_mom.o $V/$1_api.o $(ls _obj/*.o | grep -v -e lk_ssd_mom.o -e mtfb_api.o)
rm -f $V/$1_mom.o $V/$1_api.o
grep -A2 "ssd_fclk_mom_kernelILi0ELi128E" $V/$1.lk_ssd_mom.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    : //'
