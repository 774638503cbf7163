"""Rigid (self / obstacle) decisions of the CUDA path against the INDEPENDENT convex-hull oracle at scale.

The oracle (oracle/hull_oracle.py) composes its own forward kinematics from the URDF text and runs a float64 GJK on the
hulls of the collision meshes the reference scripts load; it shares nothing with the product's model compiler, packer or
kernels.  For every configuration: GPU flag (pmp_check_configs, exact hull pre-pass) vs oracle "some signed Bullet
distance <= 0", with the oracle's own clearance telling how close to the decision boundary a disagreement sits.
Run on the GPU box:   python tools/rigid_parity_report.py [n_configs]   ->  gpurun_out/rigid_parity.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from oracle import hull_oracle as H  # noqa: E402
from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
BAND = 1e-5
model = load_iiwa14()
chk = EdgeChecker(model, synthetic_worlds(1), EdgeCheckConfig(mode="ignore_all"))
scene = H.HullScene()
rng = np.random.RandomState(0)
sets = {"uniform in the joint limits": rng.uniform(model.lower, model.upper, size=(N, 7)).astype(np.float32)}
q0, q1 = synthetic_edges(model, (N + 31) // 32, seed=5)
t = (np.arange(1, 33) / 32.0)[None, :, None]
sets["steps of sweep edges"] = (q0[:, None, :] + (q1 - q0)[:, None, :] * t).reshape(-1, 7)[:N].astype(np.float32)
rep = {"band_m": BAND, "sets": {}}
for name, Q in sets.items():
    t0 = time.time()
    gpu = (chk.check_configs(torch.from_numpy(Q), detail=False).flags.cpu().numpy()[:, 0] & 1) > 0
    t_gpu = time.time() - t0
    t0 = time.time()
    ref = scene.rigid(Q.astype(np.float64))
    t_or = time.time() - t0
    mis = gpu != ref["hit"]
    clr = ref["clearance"]
    out = mis & (np.abs(clr) > BAND) & ~ref["limits"]
    rep["sets"][name] = {
        "configurations": int(len(Q)), "oracle_hits": int(ref["hit"].sum()), "gpu_hits": int(gpu.sum()),
        "mismatches": int(mis.sum()), "mismatches_outside_band": int(out.sum()),
        "max_abs_clearance_of_a_mismatch_m": float(np.abs(clr[mis]).max()) if mis.any() else 0.0,
        "configurations_inside_band": int((np.abs(clr) <= BAND).sum()),
        "seconds_gpu": t_gpu, "seconds_oracle": t_or,
    }
    print(name, rep["sets"][name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/rigid_parity.json", "w") as fp:
    json.dump(rep, fp, indent=1)
