"""How far is the stateless quasi-static plant model of the kernels from a time-stepped restatement of the reference?

Runs oracle/dynamic_oracle.py (generalized-coordinate dynamics of the plants: springs as external link torques that also
load ancestor joints, gravity, joint damping, 50 sub-steps per interpolation step, velocity-level contact solve, the arm
lagging behind its commanded configuration) along the planned path of a script scene and compares, step by step and
stem by stem, with the quasi-static oracle (== the CUDA kernels up to float32).  Variants switch single effects off to
attribute the gap.  Writes profiles/r02_dynamic_gap.json.   CPU only, a few minutes:
    python tools/dynamic_gap_report.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import edgecheck_oracle as O  # noqa: E402
from oracle.dynamic_oracle import DynamicEnv, DynParams  # noqa: E402
from plant_motion_planning_b200 import load_iiwa14  # noqa: E402
from plant_motion_planning_b200.scenarios import START, script_worlds  # noqa: E402

LIMIT = 0.30


def quasi_static(model, worlds, path, **kw):
    prm = O.OracleParams(**kw)
    Q = np.asarray(path, dtype=np.float64)
    Qg = O.lagged_configs(Q, prm) if prm.arm_lag_gain > 0 else Q
    r = O.check_configs(model, worlds, Qg, prm, Q_limits=Q)
    return r.deflection          # [n, W, P]


def dynamic(model, worlds, path, prm, settle=3000):
    env = DynamicEnv(model, worlds, prm)
    env.reset(START, settle)
    out = []
    for q in path:
        env.step(q)
        out.append(np.stack(env.deflections()))
    return np.stack(out)         # [n, W, P]


def stats(a, b):
    d = np.abs(a - b)
    ex_a, ex_b = np.any(a > LIMIT, axis=2), np.any(b > LIMIT, axis=2)
    return {"median": float(np.median(d)), "p90": float(np.percentile(d, 90)), "p99": float(np.percentile(d, 99)),
            "max": float(d.max()), "exceeded_flags_differ": float(np.mean(ex_a != ex_b)),
            "exceeded_a": float(ex_a.mean()), "exceeded_b": float(ex_b.mean())}


def main():
    model = load_iiwa14()
    with open(os.path.join(ROOT, "tests", "golden", "reference_rrt.json")) as fp:
        gold = json.load(fp)
    scene = "multi_world_ignore_all"                  # its plan ploughs through the plants (over-limit sums up to 14)
    worlds, _ = script_worlds(scene)
    path = [tuple(q) for q in gold[scene]["smoothed"]]
    path = path[::2]                                  # every second waypoint keeps the run short; steps of ~6 degrees
    report = {"scene": scene, "waypoints": len(path), "worlds": len(worlds), "limit": LIMIT, "variants": {}}
    t0 = time.time()
    qs = {"sag": quasi_static(model, worlds, path), "no_sag": quasi_static(model, worlds, path, gravity_sag=False),
          "sag_lag": quasi_static(model, worlds, path, arm_lag_gain=0.01)}
    dyn = {}
    for name, prm in (("full", DynParams()), ("no_lag", DynParams(arm_lag=False)), ("no_gravity", DynParams(gravity=False)),
                      ("no_ancestor_loading", DynParams(ancestor_loading=False)), ("friction_0.25", DynParams(friction=0.25)),
                      ("no_lag_no_gravity_no_ancestor", DynParams(arm_lag=False, gravity=False, ancestor_loading=False))):
        dyn[name] = dynamic(model, worlds, path, prm)
        print(name, "done", round(time.time() - t0, 1), "s", flush=True)
    # free-space sway of the dynamic model: the first waypoints are far from the plants
    for dn, d in dyn.items():
        for qn, q in qs.items():
            report["variants"][f"dynamic[{dn}] vs quasi_static[{qn}]"] = stats(d, q)
    report["dynamic[full] vs dynamic[no_lag]"] = stats(dyn["full"], dyn["no_lag"])
    report["dynamic[full] vs dynamic[no_gravity]"] = stats(dyn["full"], dyn["no_gravity"])
    report["dynamic[full] vs dynamic[no_ancestor_loading]"] = stats(dyn["full"], dyn["no_ancestor_loading"])
    report["note"] = ("deflections in rad, per (waypoint, world, stem), chained observer values; the dynamic model is a "
                      "restatement (oracle/dynamic_oracle.py), not PyBullet")
    out = os.path.join(ROOT, "profiles", "r02_dynamic_gap.json")
    with open(out, "w") as fp:
        json.dump(report, fp, indent=1)
    for k, v in report["variants"].items():
        print(k, {a: round(b, 4) for a, b in v.items()})
    print("wrote", out)


if __name__ == "__main__":
    main()
