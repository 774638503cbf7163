#!/bin/bash
# ncu evidence for the kernels of one sweep step (run under gpurun, one GPU):  bash tools/ncu_edge.sh <tag>
# 1) plain run must exit 0, 2) launch list with device times, 3) one --set full capture of the edge kernel
# (production and dense launch) and of the hull pre-pass.
set -u
TAG=${1:-r02}
CMD="python tools/profile_modes.py 15"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.log; exit 1; }
cat gpurun_out/ncu_plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"pmp_edge_kernel|pmp_rigid_kernel" -s 2 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -6
