"""C-ABI library: loads without a GPU, exports every symbol the header declares; table packing invariants."""
import os
import re

import numpy as np
import pytest

from plant_motion_planning_b200 import _build, _capi
from plant_motion_planning_b200.model.plants import PlantWorld, Stem, single_stem_world
from plant_motion_planning_b200.model.tables import (S_FLAGS, S_OSLOT, S_PSLOT, STEM_STRIDE, assign_slots, pack_robot,
                                                     pack_worlds)
from plant_motion_planning_b200.model.transforms import Frame
from plant_motion_planning_b200.robots import load_iiwa14
from plant_motion_planning_b200.workloads import synthetic_edges, synthetic_worlds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    with open(os.path.join(ROOT, "include", "pmp_edgecheck.h")) as fp:
        text = fp.read()
    # exported prototypes only (`int pmp_x(` / `const char* pmp_x(`); `static inline` helpers live in the header
    declared = set(re.findall(r"^(?:int|const char\*)\s+(pmp_[a-z0-9_]+)\s*\(", text, flags=re.M))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    lib = _capi.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_create_fails_loudly_without_gpu():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.load()
    ctx = C.c_void_p()
    rc = lib.pmp_create(C.byref(ctx), 0)
    assert rc != 0 and not ctx.value
    assert b"CUDA" in lib.pmp_last_error(None) or b"device" in lib.pmp_last_error(None)
    from plant_motion_planning_b200 import EdgeChecker
    with pytest.raises(RuntimeError):
        EdgeChecker(load_iiwa14(), synthetic_worlds(1))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no module of the product package may import or load it."""
    pkg = os.path.join(ROOT, "plant_motion_planning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fp:
                    src = fp.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "c_oracle" not in src, f


def test_slot_assignment_keeps_parents_alive():
    for w in synthetic_worlds(16) + synthetic_worlds(4, seed_base=50, branches=2, vert_stems=3, extensions=2):
        par, own, n = assign_slots(w)
        live = {}
        for i, s in enumerate(w.stems):
            if s.parent >= 0:
                assert live[par[i]] == s.parent, "parent frame was overwritten before its last child"
            if own[i] >= 0:
                live[own[i]] = i
        assert 1 <= n <= 4


def test_pack_worlds_layout():
    worlds = synthetic_worlds(3) + [single_stem_world([[0.4, 0.1], [0.1, -0.2]])]
    wt = pack_worlds(worlds)
    assert wt.stems.shape == (4, 6, STEM_STRIDE) and wt.stems.dtype == np.float32
    assert list(wt.n_stems) == [6, 6, 6, 2]
    iv = wt.stems.view(np.int32)
    assert (iv[3, :2, S_PSLOT] == -1).all() and (iv[3, :2, S_OSLOT] == -1).all()
    assert iv[0, 0, S_FLAGS] & 1            # plant_start on the first stem
    for w in range(4):
        R = wt.stems[w, :wt.n_stems[w], 0:9].reshape(-1, 3, 3).astype(np.float64)
        assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-6)


def test_stems_must_be_parents_first():
    bad = PlantWorld(stems=[Stem(parent=1, origin=Frame(), half_height=0.2, z_offset=0.0),
                            Stem(parent=-1, origin=Frame(), half_height=0.2, z_offset=0.0)], base_boxes=[])
    with pytest.raises(ValueError):
        pack_worlds([bad])


def test_pack_robot_masks():
    # the URDF's own 12-sphere model: floor, block, 8 block-corner points, base cylinder
    t = pack_robot(load_iiwa14(collision="urdf_spheres"))
    assert t.dof == 7 and t.n_spheres == 12 and t.n_fixed == 11
    assert sum(bin(int(m)).count("1") for m in t.self_mask) == 38
    assert all(int(t.fixed_mask[f]) == 0xFFF for f in range(10))                # obstacles: every sphere
    assert int(t.fixed_mask[10]) == 0xFFE                                       # base cylinder: not link 1 (adjacent)
    # the arm the scripts load: 18 spheres fitted to the 7 link hulls, no geometry on link 0
    m = load_iiwa14()
    t = pack_robot(m)
    assert t.dof == 7 and t.n_spheres == 18 and t.n_fixed == 10
    h = m.hulls
    assert len(h.verts) == 7 and [len(v) for v in h.verts] == [620, 773, 687, 768, 706, 688, 64]
    assert h.pairs.tolist() == [[a, b] for a in range(7) for b in range(a + 2, 7)]
    assert np.all(h.sph_radius_outer > m.sph_radius) and np.all(np.diff(h.sph_hull) >= 0)


def test_synthetic_edges_have_exact_step_count():
    from oracle import edgecheck_oracle as O
    model = load_iiwa14()
    q0, q1 = synthetic_edges(model, 512, seed=3)
    assert np.all((q1 >= model.lower) & (q1 <= model.upper))
    n = [O.num_steps(model, q0[e].astype(np.float64), q1[e].astype(np.float64), np.radians(3)) for e in range(512)]
    assert set(n) == {32}
