"""Copy the evidence of one tools/final_captures.sh run from gpurun_out/ into profiles/ and write the two small JSON
files bench.py attaches to its line (traffic.json, issue_utilization.json), keyed by the source digest of the build.
usage: python tools/publish_captures.py <tag>"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from plant_motion_planning_b200 import _build  # noqa: E402

tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
dig = _build.source_digest()
line = json.loads(open(os.path.join(G, f"bench_n1_{tag}.json")).read().strip().splitlines()[-1])
assert line["roofline"]["build_digest"] == dig, "the bench line was produced by another build than the sources in the tree"
for src, dst in ((f"bench_n1_{tag}.json", "r02_bench_n1.json"), (f"bench_ref_{tag}.json", "r02_bench_reference_arm.json"),
                 (f"launches_bench_{tag}.csv", "r02_launches_bench.csv"), (f"launches_{tag}.csv", "r02_launches_slice.csv"),
                 (f"traffic_{tag}.csv", "r02_traffic_ncu.csv")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"prof_{tag}.ncu-rep"),
                os.path.join(P, "r02_edge_rigid_ncu.csv"), "rigid_prepass_production", "edge_kernel_production",
                "rigid_prepass_dense", "edge_kernel_dense"], check=True, stdout=subprocess.DEVNULL)
if os.path.exists(os.path.join(G, f"planner_bench_{tag}.json")):          # FULL=1 captures
    shutil.copy(os.path.join(G, f"planner_bench_{tag}.json"), os.path.join(P, "r02_planner_bench.json"))
if os.path.exists(os.path.join(G, f"tree_{tag}.ncu-rep")):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, f"tree_{tag}.ncu-rep"),
                    os.path.join(P, "r02_tree_kernels_ncu.csv")], check=True, stdout=subprocess.DEVNULL)
rows = [r for r in csv.reader(open(os.path.join(G, f"traffic_{tag}.csv"))) if len(r) > 10 and r[0] == "0"]
m = {r[12]: float(r[14]) for r in rows}
E, W = 1 << 20, 64
alg = E * (2 * 7 * 4 + 16) + E * W * 8
t = {"build_digest": dig, "dram_bytes_per_launch": int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]),
     "dram_bytes_read": int(m["dram__bytes_read.sum"]), "dram_bytes_write": int(m["dram__bytes_write.sum"]),
     "gpu_time_ns": int(m["gpu__time_duration.sum"]), "lts_t_bytes": int(m["lts__t_bytes.sum"]),
     "algorithmic_bytes_per_launch": alg,
     "workload": "bench.py default: 2^20 edges x 32 steps x 64 worlds, production launch of pmp_edge_kernel<18,2,true,false,false>",
     "source": "profiles/r02_traffic_ncu.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,"
               "gpu__time_duration.sum --clock-control none, second launch, warm; tools/final_captures.sh)",
     "note": "DRAM traffic is 1.00x the algorithmic bytes (0.28 B/unit): inputs and the per-world outputs cross HBM once. L2 "
             "traffic is the register-spill write-through of the 264 B of spills per thread plus the per-warp tile traffic "
             "that misses L1."}
json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
rows = list(csv.reader(open(os.path.join(P, "r02_edge_rigid_ncu.csv"))))
h = rows[0]


def g(label, metric):
    for r in rows[2:]:
        if r[1] == label:
            return float(r[h.index(metric)])


ISS, FMA, ELI = ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
                 "smsp__warps_eligible.avg.per_cycle_active")
iu = {"build_digest": dig,
      "source": "profiles/r02_edge_rigid_ncu.csv (ncu --set full --clock-control none, 2^15-edge slice of the sweep, SM clock 1965 MHz, tools/ncu_edge.sh)",
      "metric": ISS, "production": g("edge_kernel_production", ISS) / 100, "dense": g("edge_kernel_dense", ISS) / 100,
      "fma_pipe_cycles_active": {"production": g("edge_kernel_production", FMA) / 100, "dense": g("edge_kernel_dense", FMA) / 100},
      "warps_eligible_per_scheduler": {"production": g("edge_kernel_production", ELI), "dense": g("edge_kernel_dense", ELI)},
      "hull_prepass_issue_active": g("rigid_prepass_production", ISS) / 100,
      "hull_prepass_ms_per_2^15_edges": g("rigid_prepass_production", "gpu__time_duration.sum"),
      "note": "dense: FMA-pipe bound (31 packed + 6 scalar FP32 instructions = 68 pipe cycles per sphere pair and iteration against 52 issue "
              "slots); production: issue bound on a flat profile"}
json.dump(iu, open(os.path.join(P, "issue_utilization.json"), "w"), indent=1)
r = line["roofline"]
print(json.dumps({"value": line["value"], "ms": line["ms_per_step"], "e2e": line["e2e"]["value"], "edge_ms": r["production_ms"],
                  "prepass_ms": r["hull_prepass_ms"], "dense_ms": r["dense_ms"], "frac": r["frac"],
                  "prod_req_frac": r["production_required"]["frac"], "cpu": line["cpu_baseline"]["value"],
                  "traffic": t["dram_bytes_per_launch"], "issue": iu["production"], "issue_dense": iu["dense"],
                  "prepass_ncu_ms": iu["hull_prepass_ms_per_2^15_edges"]}, indent=1))
