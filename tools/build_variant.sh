#!/bin/bash
# tools/build_variant.sh NAME "<extra nvcc flags>": builds gpurun_variants/NAME.so (a libcvmhot.so with experiment flags) for
# CVMHOT_LIB=... runs of the tools; only decode.cu / render.cu / loss.cu are rebuilt with the flags.
set -e
cd "$(dirname "$0")/../computer-vision-models_b200/csrc"
NAME=$1; shift
OUT=../../tools/variants; mkdir -p $OUT/obj_$NAME
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --fmad=true"
for f in ${FILES:-decode}; do $NVCC $FLAGS "$@" -c $f.cu -o $OUT/obj_$NAME/$f.o & done; wait
OBJS=""; for f in api loss render decode post prep; do if [ -f $OUT/obj_$NAME/$f.o ]; then OBJS="$OBJS $OUT/obj_$NAME/$f.o"; else OBJS="$OBJS build/$f.o"; fi; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/$NAME.so $OBJS -cudart static
echo built $OUT/$NAME.so
