#!/usr/bin/env python
"""Benchmark of the edge-validation hot path (BASELINE.json metric: validated edge-steps/sec, config x world).

    python bench.py [--gpus N] [--steps K] [--warmup W]          one JSON line (rank 0)
    python bench.py --impl reference ...                          the CPU arm (C twin of the oracle) on host cores

A "step" = one validate_edges pass over the synthetic sweep: per GPU 2^20 random iiwa14
edges x 32 interpolation steps x 64 sampled multi-branch plant worlds = 2.147e9 units; weak scaling (every rank
owns its own 2^20 edges, worlds replicated, one all-gather of {first_hit, n_steps, cost} per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "validated edge-steps/sec (config x world)"
UNIT = "units/s"


def flops_per_unit(K: int, P: int, A: int, W: int, D: int = 7, broad_pairs: int = 60, fixed_tests: int = 36) -> float:
    """Algorithmic FP32 work of one (configuration, world) unit as the kernel defines it (DESIGN.md section 5,
    derivation in BASELINE.md section 4): FMA = 2, add/mul/min/max/compare = 1, rcp/rsqrt/sin/cos = 1.
      per (sphere, stem, Gauss-Newton iteration), stems are the reference's BOXES ... 65
          p = R^T c - b 18, clamp to the box 6, d = p - q 3, |d|^2 6, rsqrt 1, penetration 3, lever arm 1,
          d x pf 9, Jacobian 5, weights 2, normal equations 10, sum of depths 1
      per (stem, iteration): joint rotation, frame product, box offset, 2x2 solve, sincos ... 79
      per stem: frame compose 63, final frame 49, observer 38, chain / threshold 6 ... 156
      per configuration (shared by the W worlds): FK 101 D + 18 A + 11 per broad-phase sphere pair + 34 per
          sphere-obstacle test (the exact GJK narrow phase of the hull pre-pass is not counted)"""
    per_cfg = 101.0 * D + 18.0 * A + 11.0 * broad_pairs + 34.0 * fixed_tests
    return P * (K * (65.0 * A + 79.0) + 156.0) + per_cfg / W


def flops_per_unit_survey(K: int, P: int, A: int, W: int) -> float:
    """SURVEY.md section 8(d)'s first estimate, written before any kernel existed (capsule stems, push/update model):
    F_unit(K) = K (188 P + 26 A P) + 34 P + 2066 / W."""
    return K * (188.0 * P + 26.0 * A * P) + 34.0 * P + 2066.0 / W


def flops_per_lane_required(K: int, P: int, A: int, W: int, f_touch: float, D: int = 7, broad_pairs: int = 60,
                            fixed_tests: int = 36) -> float:
    """Work the algorithm REQUIRES of one unit in the production launch: the exact contact test of every stem
    (33 per sphere + 45 per stem: p 18, max(|p| - h, 0) 6, |e|^2 5, radius test 4), and the K-step solve + observer
    only on the stems the arm touches (fraction f_touch, measured)."""
    per_cfg = 101.0 * D + 18.0 * A + 11.0 * broad_pairs + 34.0 * fixed_tests
    return P * (33.0 * A + 45.0 + f_touch * (K * (65.0 * A + 79.0) + 111.0)) + per_cfg / W


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def compare_with_oracle(got: dict, ref: dict, max_steps: int) -> dict:
    """Bit-level comparison of a GPU result with the float64 C twin on the same edges: step masks over the live
    (edge, world, step) bits, first violating step, and error percentiles of the per-(edge, world) reductions."""
    n = ref["n_steps"]
    C = (max_steps + 31) // 32
    live = np.zeros((n.shape[0], C), dtype=np.uint32)
    nc = np.minimum(n, max_steps)
    for c in range(C):
        k = np.clip(nc - 32 * c, 0, 32).astype(np.uint64)
        live[:, c] = ((np.uint64(1) << k) - np.uint64(1)).astype(np.uint32)
    bits = float(np.unpackbits(live.view(np.uint8)).sum()) * ref["rigid_mask"].shape[1]

    def mism(key):
        x = (got[key].view(np.uint32) ^ ref[key].view(np.uint32)) & live[:, None, :]
        return int(np.unpackbits(np.ascontiguousarray(x).view(np.uint8)).sum())

    def pct(err):
        e = np.abs(err).ravel()
        return {"p50": float(np.percentile(e, 50)), "p99": float(np.percentile(e, 99)),
                "p99.9": float(np.percentile(e, 99.9)), "max": float(e.max())}

    rm, xm = mism("rigid_mask"), mism("exceed_mask")
    set_r = int(np.unpackbits(np.ascontiguousarray(ref["rigid_mask"] & live[:, None, :]).view(np.uint8)).sum())
    set_x = int(np.unpackbits(np.ascontiguousarray(ref["exceed_mask"] & live[:, None, :]).view(np.uint8)).sum())
    return {
        "edges": int(n.shape[0]), "step_world_bits": int(bits),
        "n_steps_identical": bool(np.array_equal(got["n_steps"], n)),
        "rigid_bits_set_in_oracle": set_r, "rigid_bit_mismatches": rm,
        "exceed_bits_set_in_oracle": set_x, "exceed_bit_mismatches": xm,
        "mask_bit_agreement": 1.0 - (rm + xm) / max(2.0 * bits, 1.0),
        "first_hit_agreement": float(np.mean(got["first_hit"] == ref["first_hit"])),
        "max_deflection_abs_err_rad": pct(got["max_deflection"].astype(np.float64) - ref["max_deflection"]),
        "over_limit_abs_err_rad": pct(got["over_limit"].astype(np.float64) - ref["over_limit"]),
        "ee_len_abs_err_m": pct(got["ee_len"].astype(np.float64) - ref["ee_len"]),
    }


def workload(args, seed: int):
    from plant_motion_planning_b200 import load_iiwa14
    from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
    model = load_iiwa14()
    worlds = synthetic_worlds(args.worlds)
    q0, q1 = synthetic_edges(model, args.edges, seed=seed, steps=args.edge_steps)
    return model, worlds, q0, q1


def config_dict(args, n_gpus: int) -> dict:
    return {
        "workload": f"synthetic sweep: {args.edges} random iiwa14 edges x {args.edge_steps} interpolation steps x "
                    f"{args.worlds} sampled multi-branch plant worlds per GPU (BASELINE.json configs[3]), mode our_method, "
                    f"K={args.iterations} Gauss-Newton steps, deflection limit 0.30",
        "edges_per_gpu": args.edges, "steps_per_edge": args.edge_steps, "worlds": args.worlds,
        "units_per_step": int(n_gpus) * args.edges * args.edge_steps * args.worlds,
        "parallelism": f"edges sharded over {n_gpus} GPU(s), worlds replicated, all-gather of first_hit/n_steps/cost",
        "l2": "no flush: each step touches %.0f MB of inputs+outputs (> 126 MB L2); the kernel is FP32-bound, HBM traffic ~0.3 B/unit"
              % ((args.edges * 7 * 4 * 2 + args.edges * args.worlds * 8 + args.edges * 16) / 1e6),
    }


def run_reference(args, rank: int) -> None:
    """CPU arm: the reference's PyBullet path cannot run in this image, so this times the C twin of the
    oracle -- a float64 restatement of the same path -- with every host thread on a bounded sample per step."""
    if rank != 0:
        return
    from oracle.c_oracle import COracle
    model, worlds, q0, q1 = workload(argparse.Namespace(**{**vars(args), "edges": min(args.edges, 1 << 15)}), seed=0)
    co = COracle(model, worlds, iterations=args.iterations, native=True, single=True)
    cores = co.max_threads
    # bounded sample: calibrate on 256 edges, then size a step for ~args.cpu_seconds / (steps + warmup)
    t0 = time.perf_counter()
    co.validate_edges(q0[:256], q1[:256], max_steps=args.edge_steps)
    dt = max(time.perf_counter() - t0, 1e-4)
    budget = max(args.cpu_seconds / max(args.steps + args.warmup, 1), 0.25)
    n = int(min(q0.shape[0], max(256, 256 * budget / dt)))
    for _ in range(args.warmup):
        co.validate_edges(q0[:n], q1[:n], max_steps=args.edge_steps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.validate_edges(q0[:n], q1[:n], max_steps=args.edge_steps)
    el = time.perf_counter() - t0
    units = n * args.edge_steps * args.worlds
    val = units * args.steps / el
    sample = f"{n} of the sweep's edges x {args.edge_steps} steps x {args.worlds} worlds per step (seed 0)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement (oracle/edgecheck_oracle.c built in float32 with its sphere loop vectorised -- AVX2 with "
                "-march=native -fopenmp-simd --, pthreads over edges, the same per-stem early exit as the GPU path), not "
                "PyBullet: pybullet is not installable here",
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--edges", type=int, default=1 << 20, help="edges per GPU")
    ap.add_argument("--worlds", type=int, default=64)
    ap.add_argument("--edge-steps", type=int, default=32)
    ap.add_argument("--iterations", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense (roofline) leg")
    ap.add_argument("--strong", action="store_true", help="also run the config-5 strong-scaling pass at N = 1")
    ap.add_argument("--no-strong", action="store_true", help="skip the config-5 strong-scaling pass at N > 1")
    ap.add_argument("--strong-block-edges", type=int, default=1 << 20, help="edges per block of the config-5 batch (8 blocks)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker
    from plant_motion_planning_b200.distributed import all_gather_summary
    from plant_motion_planning_b200.edgecheck import EdgeBatchResult

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    model, worlds, q0, q1 = workload(args, seed=rank)
    cfg = EdgeCheckConfig(iterations=args.iterations, max_steps=args.edge_steps)
    chk = EdgeChecker(model, worlds, cfg, device=local_rank)
    E, W = args.edges, args.worlds
    a, b = torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev)
    mk = lambda shape, dt: torch.empty(shape, device=dev, dtype=dt)
    out = EdgeBatchResult(mk((E,), torch.int32), mk((E,), torch.int32), mk((E, W), torch.float32),
                          mk((E, W), torch.float32), mk((E,), torch.float32), mk((E,), torch.float32))
    gathered = mk((E * world, 3), torch.int32) if world > 1 else None

    def step():
        chk.validate_edges(a, b, out=out)
        if world > 1:
            all_gather_summary(out.first_hit, out.n_steps, out.cost, E * world, out=gathered)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = chk.launch_info()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    launches = chk.launch_info()["kernel_launches"] - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    units = world * E * args.edge_steps * W
    value = units / (ms * 1e-3)
    valid_frac = float((out.first_hit >= out.n_steps).float().mean().item())

    # ---- kernel-only timing of the dominant kernel (CUDA events recorded by the library on the launching stream,
    # around the edge kernel and around the hull pre-pass separately), production and dense ----
    chk.set_kernel_timing(True)

    def time_kernels(n_iter: int):
        chk.validate_edges(a, b, out=out)
        torch.cuda.synchronize()
        rig, edg = [], []
        for _ in range(n_iter):
            chk.validate_edges(a, b, out=out)
            r_ms, e_ms = chk.kernel_times()
            rig.append(r_ms); edg.append(e_ms)
        return float(np.mean(rig)), float(np.mean(edg))

    rigid_ms, k_ms = time_kernels(max(2, min(args.steps, 5)))
    info = chk.launch_info()
    A_, K_, P_ = model.num_spheres, args.iterations, 6
    F = flops_per_unit(K_, P_, A_, W)
    peak_tf, _ = chk.measure_fp32_peak(5)
    units_gpu = E * args.edge_steps * W
    dense_ms = None
    if not args.no_dense:
        chk.set_config(dense=True)
        _, dense_ms = time_kernels(2)
        chk.set_config(dense=False)
    chk.set_kernel_timing(False)
    # lane-level algorithmic work of the production launch: every stem needs the contact test, only touched stems
    # need the K-step solve.  The touched fraction is measured on a sample of the sweep's own configurations
    # (a stem is touched when the solve moved it away from its rest angles).
    from plant_motion_planning_b200.model.tables import S_TH0
    ns = min(E, 4096)
    mid = torch.from_numpy(0.5 * (q0[:ns] + q1[:ns])).to(dev)
    th = chk.check_configs(mid).theta
    th0 = torch.from_numpy(np.ascontiguousarray(chk.world_tables.stems[:, :, S_TH0:S_TH0 + 2])).to(dev)
    f_touch = float(((th - th0[None]).abs().sum(dim=-1) > 0).float().mean().item())
    F_lane = flops_per_lane_required(K_, P_, A_, W, f_touch)
    alg_bytes = E * (2 * 7 * 4 + 16) + E * W * 8
    # static ncu evidence (DRAM traffic, issue-slot utilisation) is attached only when it was captured for THIS build
    from plant_motion_planning_b200 import _build
    digest = _build.source_digest()

    def profile_json(name):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as fp:
                d = json.load(fp)
        except Exception:
            return None
        return d if d.get("build_digest") == digest else None

    tj = profile_json("traffic.json")
    traffic = tj.get("dram_bytes_per_launch") if tj else None
    issue = profile_json("issue_utilization.json")
    kname = f"pmp_edge_kernel<{info['n_sphere_pad']},{info['n_slot_pad']},{'true' if info['worlds_in_smem'] else 'false'},false>"
    Fs = flops_per_unit_survey(K_, P_, A_, W)
    roofline = {
        "bound": "fp32", "unit": "TFLOP/s", "peak": peak_tf,
        "peak_source": "measured on this GPU in this run: dependent-FMA probe kernel (pmp_measure_fp32_peak); "
                       "MEASURED_PEAKS.json has no FP32 CUDA-core figure (theoretical 148 SM x 128 x 2 x 1.965 GHz = 74.4)",
        "kernel": kname,
        "flops_per_unit": F,
        "flops_per_unit_survey_formula": Fs,
        "achieved": (units_gpu * F / (dense_ms * 1e-3) / 1e12) if dense_ms else None,
        "frac": (units_gpu * F / (dense_ms * 1e-3) / 1e12 / peak_tf) if dense_ms else None,
        "frac_survey_formula": (units_gpu * Fs / (dense_ms * 1e-3) / 1e12 / peak_tf) if dense_ms else None,
        "dense_ms": dense_ms,
        "note": "achieved/frac: the edge kernel alone (CUDA events around it on its stream) with dense=1 -- every unit "
                "executes all K iterations on every stem, so executed FP32 work == algorithmic work.  The production "
                "launch (dense=0) skips iterations that are provably no-ops (no lane of the warp touches the stem); "
                "results are identical.",
        "instruction_mix_ceiling": {"tflops": 60.3, "frac_of_peak": 0.83, "measured_in_this_run": False,
                                    "source": "tools/probes/mix_probe.cu on a B200 (profiles/r02_experiments.md): the pair loop's "
                                              "mix of FFMA2 / FMUL2 / FADD2 / FMNMX / MUFU on independent chains at full occupancy"},
        "production_ms": k_ms,
        "hull_prepass_ms": rigid_ms,
        "edge_kernel_share_of_step": k_ms / ms if ms > 0 else None,
        "production_required": {
            "touched_stem_fraction": f_touch, "flops_per_unit": F_lane,
            "tflops": units_gpu * F_lane / (k_ms * 1e-3) / 1e12,
            "frac": units_gpu * F_lane / (k_ms * 1e-3) / 1e12 / peak_tf,
            "note": "work the algorithm requires per lane in the production launch (contact test on every stem, solve "
                    "only on touched stems) over the edge kernel's own time; the gap to the dense fraction is SIMT "
                    "granularity: a warp runs the solve for all 32 steps when any of them touches the stem"},
        "production_equivalent_tflops": units_gpu * F / (k_ms * 1e-3) / 1e12,
        "production_equivalent_frac": units_gpu * F / (k_ms * 1e-3) / 1e12 / peak_tf,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9,
                "peak_gbs": 6534.1, "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / 6534.1,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs"},
        "traffic": traffic,
        "traffic_note": None if tj else "no ncu DRAM capture for this build (profiles/traffic.json carries another build digest)",
        "issue_slots": issue,
        "build_digest": digest,
        "launch": {k: info[k] for k in ("grid", "block", "smem_bytes", "regs_per_thread", "worlds_in_smem")},
    }

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H + sync every step ----
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    hq0, hq1 = pin((E, 7), torch.float32), pin((E, 7), torch.float32)
    hq0[:] = q0
    hq1[:] = q1
    hout = {"first_hit": pin((E,), torch.int32), "n_steps": pin((E,), torch.int32),
            "max_deflection": pin((E, W), torch.float32), "over_limit": pin((E, W), torch.float32),
            "ee_len": pin((E,), torch.float32), "cost": pin((E,), torch.float32)}
    chk.validate_edges_host(hq0, hq1, out=hout)
    barrier()
    n_e2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        chk.validate_edges_host(hq0, hq1, out=hout)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    assert np.array_equal(hout["first_hit"], out.first_hit.cpu().numpy()), "host path disagrees with device path"
    e2e = {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(hq0.nbytes + hq1.nbytes),
           "d2h_bytes_per_step": int(sum(v.nbytes for v in hout.values())),
           "api": "pmp_validate_edges_host (C ABI, pinned host buffers)"}

    # ---- BASELINE.json config 5 (scaling stress, STRONG scaling): 2^23 edges x 256 worlds sharded over the ranks,
    # all-gather of {first_hit, n_steps, cost}.  The global batch is eight blocks of 2^20 edges (seeds 1..8); rank r
    # of N owns blocks [8 r / N, 8 (r + 1) / N), so every N validates the same edges.
    strong = None
    if (world > 1 or args.strong) and not args.no_strong and 8 % world == 0:
        from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds
        W5, blocks, EB = 256, 8 // world, args.strong_block_edges
        chk5 = EdgeChecker(model, synthetic_worlds(W5), EdgeCheckConfig(iterations=args.iterations, max_steps=args.edge_steps),
                           device=local_rank)
        parts = [synthetic_edges(model, EB, seed=1 + rank * blocks + k, steps=args.edge_steps) for k in range(blocks)]
        s0 = torch.from_numpy(np.concatenate([p_[0] for p_ in parts])).to(dev)
        s1 = torch.from_numpy(np.concatenate([p_[1] for p_ in parts])).to(dev)
        E5 = s0.shape[0]
        o5 = EdgeBatchResult(mk((E5,), torch.int32), mk((E5,), torch.int32), None, None, mk((E5,), torch.float32),
                             mk((E5,), torch.float32))
        g5 = mk((E5 * world, 3), torch.int32) if world > 1 else None

        def pass5(n_edges):
            chk5.validate_edges(s0[:n_edges], s1[:n_edges], out=EdgeBatchResult(
                o5.first_hit[:n_edges], o5.n_steps[:n_edges], None, None, o5.ee_len[:n_edges], o5.cost[:n_edges]))
            if world > 1 and n_edges == E5:
                all_gather_summary(o5.first_hit, o5.n_steps, o5.cost, E5 * world, out=g5)

        pass5(min(E5, 1 << 14))
        barrier()
        s_e0, s_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_e0.record()
        pass5(E5)
        s_e1.record()
        barrier()
        t5 = torch.tensor([s_e0.elapsed_time(s_e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        ms5 = float(t5.item())
        units5 = world * E5 * args.edge_steps * W5
        i5 = chk5.launch_info()
        strong = {"scaling": "strong", "workload": f"BASELINE.json configs[4]: {world * E5} edges x {args.edge_steps} steps x {W5} worlds "
                             f"sharded over {world} GPU(s), all-gather of first_hit / n_steps / cost",
                  "edges_total": world * E5, "worlds": W5, "units": units5, "passes": 1, "ms": ms5,
                  "value": units5 / (ms5 * 1e-3), "unit": UNIT, "worlds_in_smem": bool(i5["worlds_in_smem"]),
                  "valid_edge_fraction": float((o5.first_hit >= o5.n_steps).float().mean().item())}
        chk5.close()
        del s0, s1, o5, g5

    cpu_baseline = None
    parity = None
    decision = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.c_oracle import COracle
        # the timed CPU arm: the float32 build of the C twin (same early exits as the float64 parity twin: a stem
        # that nothing touches costs one pass over the spheres), -march=native, all host threads
        co32 = COracle(model, worlds, iterations=args.iterations, native=True, single=True)
        t0 = time.perf_counter()
        co32.validate_edges(q0[:256], q1[:256], max_steps=args.edge_steps)
        dt = max(time.perf_counter() - t0, 1e-4)
        n = int(min(E, max(256, 256 * args.cpu_seconds / dt)))
        t0 = time.perf_counter()
        co32.validate_edges(q0[:n], q1[:n], max_steps=args.edge_steps)
        el = time.perf_counter() - t0
        # one core, on a sample a tenth of that (SURVEY 8d asks for both figures)
        n1 = max(64, n // (10 * max(co32.max_threads, 1)))
        t0 = time.perf_counter()
        co32.validate_edges(q0[:n1], q1[:n1], max_steps=args.edge_steps, n_threads=1)
        el1 = time.perf_counter() - t0
        # the float64 parity twin on a quarter of the sample: timed too, and compared with the GPU bit by bit
        co64 = COracle(model, worlds, iterations=args.iterations, native=True)
        n64 = max(256, min(n // 4, 8192))
        t0 = time.perf_counter()
        ref = co64.validate_edges(q0[:n64], q1[:n64], max_steps=args.edge_steps, want_masks=True)
        el64 = time.perf_counter() - t0
        cpu_baseline = {"value": n * args.edge_steps * W / el, "unit": UNIT, "cores": co32.max_threads, "kind": "port",
                        "arithmetic": "float32, sphere loop vectorised (oracle/edgecheck_oracle.c built with -DPMP_REAL=float -DPMP_SIMD_BASELINE "
                                      "-O3 -march=native -fopenmp-simd: AVX2 on this host), pthreads over edges",
                        "sample": f"first {n} edges of the same sweep x {args.edge_steps} steps x {W} worlds, {el:.1f} s",
                        "single_core_value": n1 * args.edge_steps * W / el1,
                        "single_core_sample": f"first {n1} edges, {el1:.1f} s",
                        "float64_value": n64 * args.edge_steps * W / el64,
                        "float64_sample": f"first {n64} edges, {el64:.1f} s (the parity twin)"}
        got = chk.validate_edges_host(q0[:n64], q1[:n64], want_masks=True, max_steps=args.edge_steps)
        parity = compare_with_oracle(got, ref, args.edge_steps)
        parity["sample"] = f"first {n64} edges of the sweep x {args.edge_steps} steps x {W} worlds, GPU (C ABI, host buffers) vs float64 C twin"
        # a second workload whose DECISIONS are not degenerate (on the sweep some plant of the 64 is over the limit
        # at step 0 of every edge): short ragged edges against 4 worlds, like the planner's own extension checks
        from plant_motion_planning_b200.workloads import random_edges, synthetic_worlds
        w4 = synthetic_worlds(4)
        d0, d1 = random_edges(model, 8192, seed=5, max_len=0.6)
        chk4 = EdgeChecker(model, w4, EdgeCheckConfig(iterations=args.iterations, max_steps=64), device=local_rank)
        g4 = chk4.validate_edges_host(d0, d1, want_masks=True, max_steps=64)
        r4 = COracle(model, w4, iterations=args.iterations, native=True).validate_edges(d0, d1, max_steps=64, want_masks=True)
        decision = compare_with_oracle(g4, r4, 64)
        decision["workload"] = "8192 ragged edges (1..12 steps, both ends in the joint limits) x 4 sampled worlds, limit 0.30"
        decision["valid_edge_fraction_gpu"] = float(np.mean(g4["first_hit"] >= g4["n_steps"]))
        decision["valid_edge_fraction_oracle"] = float(np.mean(r4["first_hit"] >= r4["n_steps"]))
        chk4.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "valid_edge_fraction": valid_frac, "parity_on_sample": parity, "decision_workload": decision,
        }
        if strong is not None:
            line["strong"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
