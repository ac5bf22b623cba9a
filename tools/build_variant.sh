#!/bin/bash
# Developer helper: full build of the library with extra -D flags into sift-gpu_b200/libsiftb200_<tag>.so (select with SIFT_B200_LIB).
#   tools/build_variant.sh <tag> "<defs>"
set -e
tag=$1; defs=$2
src=sift-gpu_b200/csrc; out=$src/build_$tag; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ $defs"
for u in pyramid api; do $NV -c $src/$u.cu -o $out/$u.o & done
for u in pyramid_exact detect describe match match_tc driver; do $NV --fmad=false -c $src/$u.cu -o $out/$u.o & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sift-gpu_b200/libsiftb200_$tag.so $out/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $out
