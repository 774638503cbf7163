"""Time every kernel variant under plant_motion_planning_b200/_lib/variants (GPU box).
usage: python tools/tune_run.py  -> prints ms for production and dense launches per (variant, block, ctas/SM)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
from plant_motion_planning_b200 import EdgeChecker, EdgeCheckConfig, load_iiwa14
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
m = load_iiwa14(); w = synthetic_worlds(64)
chk = EdgeChecker(m, w, EdgeCheckConfig(max_steps=32))
q0, q1 = synthetic_edges(m, 1 << 16)
a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
res = {}
for dense in (False, True):
    chk.set_config(dense=dense)
    for _ in range(2): chk.validate_edges(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): r = chk.validate_edges(a, b)
    e1.record(); torch.cuda.synchronize()
    res["dense" if dense else "prod"] = e0.elapsed_time(e1) / 3
res["info"] = chk.launch_info()
res["chk"] = int(r.first_hit.sum().item())
print(json.dumps(res))
''' % os.path.abspath(ROOT)

rows = []
libs = sorted(glob.glob(os.path.join(ROOT, "plant_motion_planning_b200/_lib/variants/*.so")))
for lib in libs:
    name = os.path.basename(lib)[7:-3]
    import re
    mm = re.match(r"t(\d+)_b(\d+)", name)
    maxt, minb = int(mm.group(1)), int(mm.group(2))
    combos = {(maxt, minb)}
    if os.environ.get("TUNE_ALL_SPLITS"):
        if maxt >= 256:
            combos.add((256, max(1, minb * maxt // 256)))
        if maxt >= 128:
            combos.add((128, max(1, minb * maxt // 128)))
    for block, per_sm in sorted(combos):
        env = dict(os.environ, PMP_LIB=lib, PMP_EDGE_BLOCK=str(block), PMP_EDGE_PER_SM=str(per_sm))
        p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        try:
            r = json.loads(p.stdout.strip().splitlines()[-1])
            rows.append((name, block, per_sm, r["prod"], r["dense"], r["info"]["regs_per_thread"], r["info"]["smem_bytes"], r["chk"]))
            print(rows[-1], flush=True)
        except Exception:
            print(name, block, per_sm, "FAILED", p.stderr[-300:], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w") as fp:
    json.dump(rows, fp)
