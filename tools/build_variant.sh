#!/bin/bash
# Build a differently tuned copy of the library: tools/build_variant.sh NAME FILE.cu "-DKNOB=.. -DKNOB2=.."
# -> orb_slam2_chinesenotes_b200/lib/variants/liborb_b200_NAME.so (FILE.cu recompiled with the knobs, every other
# object taken from the regular build).  Select it at run time with ORB_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../orb_slam2_chinesenotes_b200/csrc"
NAME=$1; SRC=$2; KNOBS=$3
make -s
mkdir -p ../lib/variants/obj
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off $KNOBS -c $SRC -o ../lib/variants/obj/${NAME}.o
OBJS=$(ls ../lib/obj/*.o | grep -v "/${SRC}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/liborb_b200_${NAME}.so $OBJS ../lib/variants/obj/${NAME}.o
echo built ../lib/variants/liborb_b200_${NAME}.so
