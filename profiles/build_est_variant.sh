#!/bin/bash
# usage: profiles/build_est_variant.sh NAME "-DEST_PROF ..."  -> scratch/variants/libNAME.so (grid_estimator.cu and mtfb_api.cu rebuilt with the flags)
set -e
cd "$(dirname "$0")/../mtf_b200/csrc"; V=../../scratch/variants; mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --compress-mode=size -fmad=false -Xcompiler -fPIC,-ffp-contract=off $2"
nvcc $FLAGS -Xptxas -v -c grid_estimator.cu -o $V/$1_est.o 2> $V/$1.est.ptxas.log &
nvcc $FLAGS -c mtfb_api.cu -o $V/$1_api.o 2> $V/$1.api.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib$1.so $V/$1_est.o $V/$1_api.o _obj/lk_ssd.o _obj/lk_ssd_f32.o _obj/lk_ssd_mom.o _obj/lk_ncc.o _obj/lk_mi.o _obj/lk_mi_aff.o _obj/pf_kernels.o _obj/pf_tracker.o _obj/preproc.o _obj/debug_kernels.o
rm -f $V/$1_est.o $V/$1_api.o
grep -E "Used" $V/$1.est.ptxas.log | head -2
