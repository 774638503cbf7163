#!/bin/bash
# Build launch-bound variants of the edge kernel (fast build: iiwa14 instantiation only) into gpurun-visible files.
# usage: bash tools/tune_variants.sh   (CPU box) ; then run tools/tune_run.py on the GPU box
set -e
cd "$(dirname "$0")/.."
mkdir -p plant_motion_planning_b200/_lib/variants
for v in "512 1" "256 2" "256 1" "192 2" "128 3" "128 4" "384 1"; do
  set -- $v
  out=plant_motion_planning_b200/_lib/variants/libpmp_t$1_b$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xptxas -v \
     -DPMP_FAST_BUILD -DPMP_EDGE_MAXT=$1 -DPMP_EDGE_MINB=$2 -o $out plant_motion_planning_b200/csrc/pmp_edgecheck.cu 2>&1 \
     | grep -A2 "edge_kernelILi12ELi2ELb1" | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/t$1 b$2: /"
  echo
done
