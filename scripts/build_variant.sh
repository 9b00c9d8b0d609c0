#!/bin/bash
# Developer helper: build an A/B variant of libblitzen_cull.so that differs from the in-tree build by -D flags on some sources.
#   usage: scripts/build_variant.sh <name> "<-D flags>" [source.cu ...]   (default source: cull_stream.cu)
# The variant lands in variants/libblitzen_cull_<name>.so (git-ignored, travels to the GPU box); select it with BLZ_CULL_LIB.
set -e
cd "$(dirname "$0")/.."
NAME=$1; DEFS=$2; shift 2
SRCS=${@:-cull_stream.cu}
mkdir -p variants/obj_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off --expt-relaxed-constexpr --extended-lambda -Xptxas -v"
OBJS=""
for f in capi.cu cull_stream.cu cull_early.cu cull_cluster.cu cull_list.cu pyramid.cu gather.cu consume.cu interop.cu raster_depth.cu; do
  if echo " $SRCS " | grep -q " $f "; then
    nvcc $FLAGS $DEFS -c blitzen_b200/csrc/$f -o variants/obj_$NAME/$f.o > variants/obj_$NAME/$f.log 2>&1 &
    OBJS="$OBJS variants/obj_$NAME/$f.o"
  else
    OBJS="$OBJS blitzen_b200/build/$f.o"
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libblitzen_cull_$NAME.so $OBJS blitzen_b200/build/blitzenCudaCull.o -cudart static -Xlinker --no-undefined
echo variants/libblitzen_cull_$NAME.so
