#!/bin/bash
# Evidence for profiles/ from ONE box (run under gpurun, one GPU):  bash tools/final_captures.sh <tag>
#  1. pytest -m gpu            2. bench.py (default flags)             3. ncu launch list of the same bench command
#  4. ncu --set full of the pre-pass + edge kernel, production and dense (2^15-edge slice)
#  5. ncu DRAM / L2 traffic of the edge kernel at the bench size (production launch)
#  6. with FULL=1: tools/planner_bench.py and the ncu capture of the tree kernels (tools/profile_tree.py)
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gputests_$TAG.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "bench launch list rc=$?"
bash tools/ncu_edge.sh $TAG > gpurun_out/ncu_edge_$TAG.log 2>&1; tail -2 gpurun_out/ncu_edge_$TAG.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
    -k regex:pmp_edge_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_$TAG.csv python tools/profile_modes.py 20 > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "traffic rc=$?"
# 6. (FULL=1) planner wall times + nearest-node throughput, ncu --set full of the tree kernels
if [ "${FULL:-0}" = "1" ]; then
  python tools/planner_bench.py > gpurun_out/planner_bench_$TAG.json 2> gpurun_out/planner_bench_$TAG.err; echo "planner rc=$?"
  python tools/profile_tree.py > gpurun_out/tree_plain_$TAG.log 2>&1 && \
    ncu --set full --clock-control none -k regex:pmp_ -s 6 -c 12 -o gpurun_out/tree_$TAG python tools/profile_tree.py > gpurun_out/tree_ncu_$TAG.log 2>&1
  echo "tree rc=$?"
fi
ls -la gpurun_out | tail -12
