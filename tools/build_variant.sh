#!/bin/bash
# tools/build_variant.sh NAME [-DEF_TRACK_THREADS=... ...] -> build_variants/libef_track_NAME.so
# Developer tool: rebuilds ef_track_kernel.cu with extra macro definitions and links it with the other (already
# built) objects, so that several tracker-kernel variants can be compared in one gpurun call (EF_TRACK_LIB=...).
set -e
cd "$(dirname "$0")/../instancefusion_b200/csrc"
name=$1; shift
mkdir -p ../../build_variants build
make -s -j4 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --ftz=true --prec-div=false --prec-sqrt=false \
     -Xcompiler -fPIC,-fvisibility=hidden "$@" -c ef_track_kernel.cu -o ../../build_variants/track_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_variants/libef_track_$name.so \
     build/ef_api.o build/ef_ops_image.o build/ef_ops_reduce.o build/ef_ops_depth.o build/ef_ops_predict.o build/ef_build_fused.o build/ef_track_dispatch.o build/ef_track_kernel_t384.o ../../build_variants/track_$name.o
rm -f ../../build_variants/track_$name.o
echo built build_variants/libef_track_$name.so
