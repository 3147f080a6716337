#!/bin/bash
# liborb_b200_checked.so: the kernels added in the last session of round 2 (orb_match_batch.cu, orb_match_bow.cu, orb_frame.cu)
# recompiled with -DORB_BOUNDS_CHECK (index checks that print and trap), every other object from the regular build.
# Run the GPU tests against it with ORB_B200_LIB=orb_slam2_chinesenotes_b200/lib/variants/liborb_b200_checked.so
set -e
cd "$(dirname "$0")/../orb_slam2_chinesenotes_b200/csrc"
make -s
mkdir -p ../lib/variants/obj
OBJS=$(ls ../lib/obj/*.o)
for f in orb_match_batch orb_match_bow orb_frame; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -DORB_BOUNDS_CHECK -c $f.cu -o ../lib/variants/obj/${f}_checked.o
  OBJS=$(echo "$OBJS" | grep -v "/${f}.cu.o")
  OBJS="$OBJS ../lib/variants/obj/${f}_checked.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/liborb_b200_checked.so $OBJS
echo built ../lib/variants/liborb_b200_checked.so
