"""GPU-vs-oracle error distributions (run on the GPU box): python tools/parity_report.py [n_configs]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

from oracle import edgecheck_oracle as O  # noqa: E402
from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import random_edges, synthetic_edges, synthetic_worlds  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = 8
model = load_iiwa14()
worlds = synthetic_worlds(W)
cfg = EdgeCheckConfig()
chk = EdgeChecker(model, worlds, cfg)
rng = np.random.RandomState(0)
Q = rng.uniform(model.lower, model.upper, size=(B, 7)).astype(np.float32)
rep = {}

pos, quat, com = chk.fk(torch.from_numpy(Q))
opos, oquat, ocom = O.link_poses(model, Q.astype(np.float64))
rep["fk_pos_max_err"] = float(np.abs(pos.cpu().numpy() - opos).max())
rep["fk_com_max_err"] = float(np.abs(com.cpu().numpy() - ocom).max())
qg = quat.cpu().numpy()
sgn = np.sign(np.sum(qg * oquat, axis=-1, keepdims=True))
rep["fk_quat_max_err"] = float(np.abs(qg * sgn - oquat).max())

for mode in ("our_method", "avoid_all", "ignore_all"):
    chk.set_config(mode=mode)
    prm = O.OracleParams(mode={"our_method": 0, "avoid_all": 1, "ignore_all": 2}[mode], reg_eps=cfg.reg_eps)
    t0 = time.time()
    res = O.check_configs(model, worlds, Q.astype(np.float64), prm)
    t_or = time.time() - t0
    g = chk.check_configs(torch.from_numpy(Q))
    torch.cuda.synchronize()
    fl = g.flags.cpu().numpy()
    gr, gx = (fl & 1) > 0, (fl & 2) > 0
    dd = np.abs(g.deflection.cpu().numpy() - res.deflection)
    dt = np.abs(g.theta.cpu().numpy() - res.theta)
    rep[mode] = {
        "oracle_s": t_or,
        "rigid_mismatch": int((gr != res.rigid).sum()), "rigid_true": int(res.rigid.sum()),
        "exceeded_mismatch": int((gx != res.exceeded).sum()), "exceeded_true": int(res.exceeded.sum()),
        "defl_err_pct_50_90_99_999_max": [float(v) for v in np.percentile(dd, [50, 90, 99, 99.9, 100])],
        "theta_err_pct_50_90_99_999_max": [float(v) for v in np.percentile(dt, [50, 90, 99, 99.9, 100])],
        "frac_defl_err_gt_1e-4": float((dd > 1e-4).mean()), "frac_defl_err_gt_1e-3": float((dd > 1e-3).mean()),
        "ee_max_err": float(np.abs(g.ee.cpu().numpy() - res.ee).max()),
    }
    if mode == "our_method":
        touched = np.abs(res.theta).sum(axis=-1) > 0
        rep[mode]["touched_frac"] = float(touched.mean())
        rep[mode]["defl_err_touched_pct_50_90_99"] = [float(v) for v in np.percentile(dd[touched], [50, 90, 99])]
        bad = np.argwhere(dd > 1e-3)
        rep[mode]["worst_examples"] = [
            {"b": int(b), "w": int(w), "s": int(s), "gpu": float(g.deflection[b, w, s]), "oracle": float(res.deflection[b, w, s]),
             "theta_gpu": g.theta[b, w, s].cpu().tolist(), "theta_or": res.theta[b, w, s].tolist()} for b, w, s in bad[:8]]

# edges
chk.set_config(mode="our_method")
q0, q1 = random_edges(model, 256, seed=3)
out = chk.validate_edges(torch.from_numpy(q0), torch.from_numpy(q1), want_masks=True)
torch.cuda.synchronize()
prm = O.OracleParams(reg_eps=cfg.reg_eps)
nmis = fmis = 0
errs = {"max_defl": 0.0, "over": 0.0, "ee_len": 0.0, "cost": 0.0}
for e in range(64):
    r = O.validate_edge(model, worlds, q0[e].astype(np.float64), q1[e].astype(np.float64), prm)
    nmis += int(r.n_steps != int(out.n_steps[e]))
    fmis += int(r.first_hit != int(out.first_hit[e]))
    errs["max_defl"] = max(errs["max_defl"], float(np.abs(out.max_deflection[e].cpu().numpy() - r.max_deflection).max()))
    errs["over"] = max(errs["over"], float(np.abs(out.over_limit[e].cpu().numpy() - r.over_limit).max()))
    errs["ee_len"] = max(errs["ee_len"], abs(float(out.ee_len[e]) - r.ee_len))
    errs["cost"] = max(errs["cost"], abs(float(out.cost[e]) - r.cost))
rep["edges"] = {"n_steps_mismatch": nmis, "first_hit_mismatch": fmis, **errs, "launch": chk.launch_info()}

# throughput quick look
q0, q1 = synthetic_edges(model, 1 << 14)
worlds64 = synthetic_worlds(64)
chk64 = EdgeChecker(model, worlds64, cfg)
a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
for dense in (False, True):
    chk64.set_config(dense=dense)
    for _ in range(2):
        chk64.validate_edges(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = chk64.validate_edges(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rep["throughput_dense" if dense else "throughput"] = {
        "ms": ms, "units_per_s": (1 << 14) * 32 * 64 / (ms * 1e-3), "valid_frac": float(r.valid.float().mean()),
        "launch": chk64.launch_info()}
tf, ms = chk64.measure_fp32_peak()
rep["fp32_peak_tflops"] = tf
print(json.dumps(rep, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/parity_report.json", "w") as fp:
    json.dump(rep, fp, indent=1)
