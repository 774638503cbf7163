"""Fit the arm's sphere model to the convex hulls the reference scripts load.   Run in the build container.

Input : tests/golden/iiwa14_collision_hulls.npz (tools/extract_iiwa_hulls.py; hull of every collision mesh of
        iiwa14_polytope_collision_edit.urdf).  Target body per link T = hull (+) 1 mm Bullet margin.
Output: plant_motion_planning_b200/data/iiwa14_hull_spheres_fit.json: per link k spheres (centre, radius, inscribed
        radius = depth of the centre + margin) and the measured two-sided deviation of their union from T.

Method (set cover, no local minima): candidate centres are a grid inside the hull; a candidate with protrusion
allowance e is the sphere of radius depth(c) + margin + e (it sticks out of T by at most e).  A point of T's surface
counts as covered when it lies within e of such a sphere.  For a given e a greedy set cover picks spheres; the
smallest e for which k spheres cover every surface sample is found by bisection.  Sphere counts are handed out to the
link with the largest e until the budget is used.  Deviations are then re-measured on denser samples.
    python tools/fit_arm_spheres.py --budget 32
"""
import argparse
import json
import os
import sys

import numpy as np
from scipy.spatial import ConvexHull

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "..", "tests", "golden", "iiwa14_collision_hulls.npz")
MARGIN = 1.0e-3


def fib_dirs(n):
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


class Link:
    def __init__(self, V, faces, n_surf=4000, grid=0.008, seed=0):
        self.V = V
        hull = ConvexHull(V)
        self.N = hull.equations[:, :3]
        self.D = -hull.equations[:, 3]           # inside: N x <= D
        self.surf = self.surface_samples(faces, n_surf, seed)
        lo, hi = V.min(0), V.max(0)
        ax = [np.arange(lo[k] + grid / 2, hi[k], grid) for k in range(3)]
        G = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
        dep = -self.depth(G)
        keep = dep > 0.01                         # centres at least 1 cm inside
        self.cand = G[keep]
        self.cand_depth = dep[keep]

    def surface_samples(self, faces, n, seed):
        V = self.V
        tri = V[faces]
        nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        area = 0.5 * np.linalg.norm(nrm, axis=1)
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
        rng = np.random.default_rng(seed)
        f = rng.choice(len(tri), size=n, p=area / area.sum())
        u, v = rng.random(n), rng.random(n)
        flip = u + v > 1
        u[flip], v[flip] = 1 - u[flip], 1 - v[flip]
        P = tri[f, 0] + u[:, None] * (tri[f, 1] - tri[f, 0]) + v[:, None] * (tri[f, 2] - tri[f, 0])
        out = V - V.mean(0)
        return np.concatenate([P + MARGIN * nrm[f], V + MARGIN * out / np.linalg.norm(out, axis=1, keepdims=True)])

    def depth(self, X):
        """signed distance to the hull by planes: < 0 inside (exact), > 0 outside (lower bound)."""
        return np.max(X @ self.N.T - self.D[None], axis=1)

    def cover(self, e, kmax):
        """greedy set cover with allowance e: returns chosen candidate indices, or None if more than kmax are needed."""
        dist = np.linalg.norm(self.surf[:, None, :] - self.cand[None], axis=2)       # [S, C]
        cov = dist <= (self.cand_depth + MARGIN + 2 * e)[None]                      # protrude e, under-cover e
        left = np.ones(len(self.surf), dtype=bool)
        chosen = []
        while left.any():
            if len(chosen) >= kmax:
                return None
            gain = cov[left].sum(axis=0)
            j = int(np.argmax(gain))
            if gain[j] == 0:
                return None
            chosen.append(j)
            left &= ~cov[:, j]
        return chosen

    def fit(self, k):
        lo, hi = 0.0, 0.08
        best = None
        for _ in range(12):
            mid = 0.5 * (lo + hi)
            ch = self.cover(mid, k)
            if ch is None:
                lo = mid
            else:
                hi, best = mid, (mid, ch)
        e, ch = best
        c = self.cand[ch]
        r = self.cand_depth[ch] + MARGIN + e
        return c, r, e

    def deviations(self, c, r, faces):
        surf = self.surface_samples(faces, 20000, 7)
        e_in = (np.linalg.norm(surf[:, None, :] - c[None], axis=2) - r[None]).min(axis=1)
        dirs = fib_dirs(2000)
        S = (c[:, None, :] + r[:, None, None] * dirs[None]).reshape(-1, 3)
        own = np.repeat(np.arange(len(r)), len(dirs))
        dd = np.linalg.norm(S[:, None, :] - c[None], axis=2) - r[None]
        dd[np.arange(len(S)), own] = 1.0
        e_out = (self.depth(S) - MARGIN)[dd.min(axis=1) > 0]
        return float(max(e_in.max(), 0.0)), float(max(e_out.max(), 0.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=int, default=32)
    ap.add_argument("--kmax", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(HERE, "..", "plant_motion_planning_b200", "data", "iiwa14_hull_spheres_fit.json"))
    args = ap.parse_args()
    d = np.load(FIX)
    names = [str(n) for n in d["link_names"]]
    links = [Link(d[f"verts_{n}"], d[f"faces_{n}"]) for n in names]
    table = {}
    for n, l in zip(names, links):
        for k in range(1, args.kmax + 1):
            table[(n, k)] = l.fit(k)
        print(n, len(l.cand), "candidates; allowance e(k) mm:", [round(1e3 * table[(n, k)][2], 1) for k in range(1, args.kmax + 1)], flush=True)
    ks = [1] * len(names)
    while sum(ks) < args.budget:
        err = [table[(n, k)][2] if k < args.kmax else -1.0 for n, k in zip(names, ks)]
        i = int(np.argmax(err))
        if err[i] < 0:
            break
        ks[i] += 1
    print("allocation", ks)
    out = {"margin": MARGIN, "links": []}
    for n, l, k in zip(names, links, ks):
        c, r, e = table[(n, k)]
        order = np.argsort(c @ np.linalg.svd(l.V - l.V.mean(0), full_matrices=False)[2][0])
        c, r = c[order], r[order]
        ein, eout = l.deviations(c, r, d[f"faces_{n}"])
        r_in = np.maximum(-l.depth(c) + MARGIN, 0.0)
        out["links"].append({"link": n, "center": c.tolist(), "radius": r.tolist(), "radius_inscribed": r_in.tolist(),
                             "under_cover_m": ein, "over_cover_m": eout})
        print(n, k, "spheres: under-cover %.2f mm  over-cover %.2f mm" % (1e3 * ein, 1e3 * eout))
    with open(args.out, "w") as fp:
        json.dump(out, fp, indent=1)
    print("wrote", os.path.abspath(args.out))


if __name__ == "__main__":
    sys.exit(main())
