"""Random plant topologies for stress tests (arbitrary parents-first trees, both observer kinds).  Test infrastructure."""
import numpy as np

from plant_motion_planning_b200.model.plants import OBS_ANGLE, OBS_MOD3, PlantWorld, Stem
from plant_motion_planning_b200.model.tables import assign_slots
from plant_motion_planning_b200.model.transforms import Frame, rpy_to_matrix


def random_world(seed: int, n_stems: int, max_slots: int = 4) -> PlantWorld:
    rng = np.random.RandomState(seed)
    while True:
        stems = []
        base_xy = np.array([rng.uniform(0.0, 0.35), rng.uniform(-0.5, 0.0)])
        obs = OBS_MOD3 if rng.rand() < 0.5 else OBS_ANGLE
        for k in range(n_stems):
            # parents-first, DFS-ish: attach to one of the last few stems so that the live-frame count stays small
            parent = -1 if k == 0 or rng.rand() < 0.1 else int(rng.randint(max(0, k - 3), k))
            h = rng.uniform(0.12, 0.22)
            if parent < 0:
                origin = Frame(rpy_to_matrix((0, 0, rng.uniform(-1, 1))),
                               np.array([base_xy[0] + rng.uniform(-0.1, 0.1), base_xy[1] + rng.uniform(-0.1, 0.1), 0.05]))
            else:
                ph = stems[parent].half_height
                origin = Frame(rpy_to_matrix(rng.uniform(-1.2, 1.2, 3)),
                               np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.05, 2 * ph)]))
            s = Stem(parent=parent, origin=origin, half_height=h, z_offset=rng.choice([0.0, 0.05]),
                     plant_start=(parent < 0), is_main=bool(rng.rand() < 0.4), obs_kind=obs)
            stems.append(s)
        world = PlantWorld(stems=stems, base_boxes=[(Frame(np.eye(3), np.array([base_xy[0], base_xy[1], 0.025])),
                                                     np.array([0.03, 0.03, 0.025]))])
        if obs == OBS_ANGLE:
            from plant_motion_planning_b200.model.tables import stem_rest_frames
            rest = stem_rest_frames(world)
            for s, fr in zip(world.stems, rest):
                s.obs_ref_point = np.array([base_xy[0], base_xy[1], float(fr.t[2])])   # like base_points of the pickle
        if assign_slots(world)[2] <= max_slots:
            return world
        seed += 7919
        rng = np.random.RandomState(seed)
