#!/bin/bash
# Experiment builds of the library: bash tools/exp_variants.sh TAG "-DFLAG ..." [TAG2 "..."]  ->  _lib/variants/libpmp_TAG.so
# Only pmp_edgecheck.cu and the A = 18 instantiations see the flags; the other translation units come from a base build.
set -e
cd "$(dirname "$0")/.."
C=plant_motion_planning_b200/csrc
OUT=plant_motion_planning_b200/_lib/variants
BASE=$OUT/obj          # scratch objects next to the variant libraries (git-ignored; delete _lib/variants before a gpurun snapshot you do not need them in)
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p $OUT $BASE
for f in pmp_inst_a12 pmp_inst_a16 pmp_inst_a20 pmp_inst_a32 pmp_inst_a12f pmp_inst_a16f pmp_inst_a20f pmp_inst_a32f; do
  if [ ! -f $BASE/$f.o ] || [ $C/pmp_kernels.cuh -nt $BASE/$f.o ]; then nvcc $FLAGS -c -o $BASE/$f.o $C/$f.cu > $BASE/$f.log 2>&1 & fi
done
wait
while [ $# -ge 2 ]; do
  tag=$1; extra=$2; shift 2
  for f in pmp_edgecheck pmp_inst_a18 pmp_inst_a18f; do
    nvcc $FLAGS $extra -c -o $BASE/${f}_$tag.o $C/$f.cu > $BASE/${f}_$tag.log 2>&1 &
  done
  wait
  nvcc --shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libpmp_$tag.so $BASE/pmp_inst_a12.o $BASE/pmp_inst_a16.o $BASE/pmp_inst_a20.o $BASE/pmp_inst_a32.o \
     $BASE/pmp_inst_a12f.o $BASE/pmp_inst_a16f.o $BASE/pmp_inst_a20f.o $BASE/pmp_inst_a32f.o $BASE/pmp_edgecheck_$tag.o $BASE/pmp_inst_a18_$tag.o $BASE/pmp_inst_a18f_$tag.o
  echo -n "$tag: "; grep -A2 "edge_kernelILi18ELi2ELb1ELb0" $BASE/pmp_inst_a18_$tag.log | grep -E "registers|spill" | tr '\n' ' '; echo
done
