"""Time breakdown of one sweep slice: hull pre-pass on/off, sphere-count variants, production vs dense.
usage: python tools/profile_variants.py [log2_edges]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from plant_motion_planning_b200 import EdgeCheckConfig, EdgeChecker, load_iiwa14  # noqa: E402
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 15)
worlds = synthetic_worlds(64)


def run(tag, model, **cfg):
    chk = EdgeChecker(model, worlds, EdgeCheckConfig(max_steps=32, **cfg))
    q0, q1 = synthetic_edges(model, n)
    a, b = torch.from_numpy(q0).cuda(), torch.from_numpy(q1).cuda()
    for dense in (False, True):
        chk.set_config(dense=dense)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        chk.validate_edges(a, b)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            chk.validate_edges(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{tag:28s} {'dense' if dense else 'prod '} {ms:9.3f} ms  {n * 32 * 64 / ms / 1e6:8.3f} G units/s", chk.launch_info()["regs_per_thread"], flush=True)
    chk.close()


hull = load_iiwa14()
run("hulls+18 spheres", hull)
run("18 spheres, no pre-pass", hull, use_hulls=False)
run("18 spheres, no sag", hull, use_hulls=False, gravity_sag=False)
run("urdf 12 spheres", load_iiwa14(collision="urdf_spheres"))
run("urdf 12 spheres, no sag", load_iiwa14(collision="urdf_spheres"), gravity_sag=False)
