#!/bin/bash
# ncu evidence for the edge kernel (run under gpurun, one GPU):  bash tools/ncu_edge.sh <tag>
# 1) plain run must exit 0, 2) launch list with device times, 3) one --set full capture of the edge kernel.
set -u
TAG=${1:-r01}
CMD="python bench.py --edges 131072 --steps 2 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pmp_edge_kernel -s 1 -c 2 -o gpurun_out/prof_$TAG $CMD --no-dense > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
