#!/bin/bash
# tools/build_batch_variant.sh NAME -DEF_TRACK_GROUPS=2 -DEF_TRACK_THREADS=192 [...] -> build_variants/libef_track_NAME.so
# Developer tool: the product library with another shape of the BATCHED tracker kernel (EF_TRACK_LIB=... selects it).
set -e
cd "$(dirname "$0")/../instancefusion_b200/csrc"
name=$1; shift
mkdir -p ../../build_variants build
make -s -j4 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --ftz=true --prec-div=false --prec-sqrt=false \
     -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v "$@" -c ef_track_kernel.cu -o ../../build_variants/batch_$name.o 2> ../../build_variants/batch_$name.log
grep -A2 "k_trackILb0" ../../build_variants/batch_$name.log | grep -E "spill|Used" | tr '\n' ' '; echo
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_variants/libef_track_$name.so \
     build/ef_api.o build/ef_ops_image.o build/ef_ops_reduce.o build/ef_ops_depth.o build/ef_ops_predict.o build/ef_build_fused.o build/ef_track_dispatch.o \
     build/ef_track_kernel_t256.o build/ef_track_kernel_t384.o ../../build_variants/batch_$name.o
rm -f ../../build_variants/batch_$name.o
echo built build_variants/libef_track_$name.so
