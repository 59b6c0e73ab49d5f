#!/bin/bash
# Build a tuning variant of the library: tools/build_variant.sh NAME "-DBLP_UNROLL_CC=4 ..." [file.cu ...]
# -> blp_b200/variants/libblp_b200_NAME.so (select with BLP_B200_LIB=...).  Only the listed sources (default:
# blp_sweep.cu) are recompiled with the extra flags; everything else links from the normal objects.
set -eu
NAME=$1; FLAGS=$2; shift 2
SRCS=${@:-blp_sweep.cu}
cd "$(dirname "$0")/../blp_b200/csrc"
make -s -j8
mkdir -p ../variants/obj_$NAME
OBJS=""
for f in *.cu; do
  o=${f%.cu}.o
  if echo " $SRCS " | grep -q " $f "; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $FLAGS -c -o ../variants/obj_$NAME/$o $f
    OBJS="$OBJS ../variants/obj_$NAME/$o"
  else
    OBJS="$OBJS $o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../variants/libblp_b200_$NAME.so $OBJS
echo "built blp_b200/variants/libblp_b200_$NAME.so"
