#!/bin/bash
# Build launch-bound variants of the edge kernel into plant_motion_planning_b200/_lib/variants/ (they travel to the GPU box).
# usage: bash tools/tune_variants.sh "512 1" "256 2" ...   then on the GPU box: python tools/tune_run.py
# Extra nvcc defines can be passed through EXTRA, e.g. EXTRA="-DPMP_SOMETHING" bash tools/tune_variants.sh "512 1"
set -e
cd "$(dirname "$0")/.."
C=plant_motion_planning_b200/csrc
OUT=plant_motion_planning_b200/_lib/variants
mkdir -p $OUT/obj
[ $# -eq 0 ] && set -- "512 1" "256 2" "256 1" "640 1"
for v in "$@"; do
  set -- $v
  tag=t$1_b$2${TAG:-}
  FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -DPMP_EDGE_MAXT=$1 -DPMP_EDGE_MINB=$2 ${EXTRA:-}"
  for f in pmp_edgecheck pmp_inst_a12 pmp_inst_a16 pmp_inst_a32 pmp_inst_a12f pmp_inst_a16f pmp_inst_a32f; do
    nvcc $FLAGS -c -o $OUT/obj/${f}_$tag.o $C/$f.cu > $OUT/obj/${f}_$tag.log 2>&1 &
  done
  wait
  nvcc --shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libpmp_$tag.so $OUT/obj/*_$tag.o
  grep -A2 "edge_kernelILi12ELi2ELb1" $OUT/obj/pmp_inst_a12_$tag.log | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/$tag: /"
  echo
done
rm -rf $OUT/obj
