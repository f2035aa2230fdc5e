#!/bin/bash
# tools/build_variant.sh NAME [-DEF_X_...=1 ...] -> build_variants/libef_track_NAME.so
# Developer tool: rebuilds the 256-thread single-launch tracker kernel with extra macro definitions and links it with the
# other (already built) objects, so that several variants can be compared in one gpurun call (EF_TRACK_LIB=...).
# VARIANT_THREADS=384 rebuilds the 384-thread variant instead; SHAPE_THREADS=320 puts another CTA shape into that slot.
set -e
cd "$(dirname "$0")/../instancefusion_b200/csrc"
name=$1; shift
T=${VARIANT_THREADS:-256}
other=$([ "$T" = 256 ] && echo 384 || echo 256)
mkdir -p ../../build_variants build
make -s -j4 >/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --ftz=true --prec-div=false --prec-sqrt=false \
     -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -DEF_TRACK_THREADS=${SHAPE_THREADS:-$T} -DEF_TRACK_NAME_THREADS=$T "$@" -c ef_track_kernel.cu -o ../../build_variants/track_$name.o 2> ../../build_variants/track_$name.log
echo "$name: $(grep -A2 'k_trackILb0' ../../build_variants/track_$name.log | grep -E 'spill|Used' | sed 's/ptxas info    : //' | tr '\n' ' ')"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_variants/libef_track_$name.so \
     build/ef_api.o build/ef_ops_image.o build/ef_ops_reduce.o build/ef_ops_depth.o build/ef_ops_predict.o build/ef_build_fused.o build/ef_track_dispatch.o \
     build/ef_track_kernel_t$other.o build/ef_track_kernel_batch.o ../../build_variants/track_$name.o
rm -f ../../build_variants/track_$name.o
