"""bench.py's reference arm runs anywhere (it times the CPU port): check the JSON line's contract on a tiny sample,
and that the native arm refuses to run without a GPU instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--edges", "1024", "--worlds", "4", "--cpu-seconds", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"].startswith("validated edge-steps/sec") and d["unit"] == "units/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--edges", "64",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_parity_block_counts_bits_over_live_steps_only():
    """bench.py's parity_on_sample / decision_workload blocks: mismatches are counted over the live (edge, world, step) bits
    (steps beyond an edge's n_steps are not evidence), percentiles over the per-(edge, world) reductions."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    E, W, C = 3, 2, 2
    n = np.array([5, 40, 64], dtype=np.int32)
    ref = {"n_steps": n, "first_hit": np.array([5, 7, 64], np.int32),
           "rigid_mask": np.zeros((E, W, C), np.uint32), "exceed_mask": np.zeros((E, W, C), np.uint32),
           "max_deflection": np.zeros((E, W)), "over_limit": np.zeros((E, W)), "ee_len": np.ones(E)}
    ref["rigid_mask"][1, 0, 0] = 1 << 7
    got = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ref.items()}
    got["max_deflection"] = got["max_deflection"].astype(np.float32)
    got["over_limit"] = got["over_limit"].astype(np.float32)
    got["ee_len"] = got["ee_len"].astype(np.float32)
    got["rigid_mask"][0, 1, 0] |= 1 << 9           # beyond edge 0's 5 steps: not counted
    got["exceed_mask"][1, 1, 1] |= 1 << 3          # step 35 of edge 1 (40 steps): counted
    got["exceed_mask"][1, 1, 1] |= 1 << 20         # step 52 of edge 1: beyond its steps, not counted
    got["first_hit"][2] = 10
    got["max_deflection"][2, 1] = 0.25
    r = bench.compare_with_oracle(got, ref, 64)
    assert r["step_world_bits"] == (5 + 40 + 64) * W and r["n_steps_identical"]
    assert r["rigid_bits_set_in_oracle"] == 1 and r["rigid_bit_mismatches"] == 0
    assert r["exceed_bit_mismatches"] == 1
    assert abs(r["mask_bit_agreement"] - (1 - 1 / (2 * 218))) < 1e-12
    assert abs(r["first_hit_agreement"] - 2 / 3) < 1e-12
    assert r["max_deflection_abs_err_rad"]["max"] == 0.25 and r["ee_len_abs_err_m"]["max"] == 0.0
