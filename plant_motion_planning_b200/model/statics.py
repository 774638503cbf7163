"""Rest pose of a plant under gravity and the reference's stem springs (host-side set-up, float64).

In the reference a plant is created with every joint at 0 and the deflection observers capture their reference pose
right then (representation.py:206-212), before the first simulation step.  Then gravity (-9.8, stem mass 10 kg, centre
of mass 0.25 m up the stem: rand_gen_plants_programmatically_main.py:255-266) pulls every tilted branch down until the
springs of env.py:127-144 balance it, and the `_mod3` observer reports that sag as deflection from then on.  The
springs are EXTERNAL LINK TORQUES [-kp tx, -kp ty, 0] in the stem link's frame (kp = 500), so the torque applied to a
stem also loads every ancestor joint (generalized force = joint axis . torque), with no reaction on the parent link.

This module solves that static balance once per world:
    for every joint j (axis a_j through o_j):   sum over stems s in the subtree of j of
        a_j . ( R_s (-kp tx_s, -kp ty_s, 0)  +  (com_s - o_j) x (0, 0, -m g) )  =  0
by Newton's method with a finite-difference Jacobian, vectorised over a batch of worlds that share one topology.
The edge-validation kernels start every stem's contact solve from this pose and report its observer value for stems
that nothing touches (stem-record fields PMP_S_TH0 / PMP_S_RAW0).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

SPRING_KP = 500.0      # env.py:128
STEM_MASS = 10.0       # rand_gen_plants_programmatically_main.py:255
GRAVITY = 9.8          # pybullet_tools/utils.py:1218-1221
REST_TOL = 1e-7        # residual (relative to kp) a rest pose must reach
REST_MAX_ANGLE = 0.75  # rad: beyond this the "rest pose" is a buckled plant, not a sag


def _mm(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """[..., 3, 3] x [..., 3, 3] with a fixed summation order (no BLAS: the result of a world must not depend on the
    size of the batch it is packed in)."""
    return (A[..., :, 0, None] * B[..., 0, None, :] + A[..., :, 1, None] * B[..., 1, None, :]) + A[..., :, 2, None] * B[..., 2, None, :]


def _mv(A: np.ndarray, v: np.ndarray) -> np.ndarray:
    return (A[..., :, 0] * v[..., 0, None] + A[..., :, 1] * v[..., 1, None]) + A[..., :, 2] * v[..., 2, None]


def _mtv(A: np.ndarray, v: np.ndarray) -> np.ndarray:
    return (A[..., 0, :] * v[..., 0, None] + A[..., 1, :] * v[..., 1, None]) + A[..., 2, :] * v[..., 2, None]


def _dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _rx_ry(tx: np.ndarray, ty: np.ndarray) -> np.ndarray:
    cx, sx, cy, sy = np.cos(tx), np.sin(tx), np.cos(ty), np.sin(ty)
    M = np.zeros(tx.shape + (3, 3))
    M[..., 0, 0], M[..., 0, 2] = cy, sy
    M[..., 1, 0], M[..., 1, 1], M[..., 1, 2] = sx * sy, cx, -sx * cy
    M[..., 2, 0], M[..., 2, 1], M[..., 2, 2] = -cx * sy, sx, cx * cy
    return M


def _rx(tx: np.ndarray) -> np.ndarray:
    cx, sx = np.cos(tx), np.sin(tx)
    M = np.zeros(tx.shape + (3, 3))
    M[..., 0, 0] = 1.0
    M[..., 1, 1], M[..., 1, 2] = cx, -sx
    M[..., 2, 1], M[..., 2, 2] = sx, cx
    return M


def static_residual(parent: Sequence[int], origin_R: np.ndarray, origin_t: np.ndarray, com_z: np.ndarray,
                    theta: np.ndarray, kp: float, mass: np.ndarray, g: float) -> np.ndarray:
    """Generalized force of springs + gravity on every joint, [W, 2P].  origin_R [W,P,3,3], origin_t [W,P,3],
    com_z [W,P], theta [W,P,2]; `parent[s]` < s or -1 (same topology for the whole batch)."""
    W, P = theta.shape[0], len(parent)
    R = [None] * P; t = [None] * P; ax = [None] * P; ay = [None] * P
    for s in range(P):
        if parent[s] < 0:
            Rv = origin_R[:, s]
            t[s] = origin_t[:, s]
        else:
            Rp = R[parent[s]]
            Rv = _mm(Rp, origin_R[:, s])
            t[s] = _mv(Rp, origin_t[:, s]) + t[parent[s]]
        Rx = _mm(Rv, _rx(theta[:, s, 0]))
        ax[s] = Rv[:, :, 0]
        ay[s] = Rx[:, :, 1]
        R[s] = _mm(Rv, _rx_ry(theta[:, s, 0], theta[:, s, 1]))
    F = np.zeros((W, 2 * P))
    for s in range(P):
        tl = np.stack([-kp * theta[:, s, 0], -kp * theta[:, s, 1], np.zeros(W)], axis=1)
        tw = _mv(R[s], tl)
        com = R[s][:, :, 2] * com_z[:, s, None] + t[s]
        fg = np.zeros((W, 3)); fg[:, 2] = -mass[s] * g
        a = s
        while a >= 0:
            tot = tw + np.cross(com - t[a], fg)
            F[:, 2 * a] += _dot(ax[a], tot)
            F[:, 2 * a + 1] += _dot(ay[a], tot)
            a = parent[a]
    return F


def rest_pose_batch(parent: Sequence[int], origin_R: np.ndarray, origin_t: np.ndarray, com_z: np.ndarray,
                    kp: float = SPRING_KP, mass=STEM_MASS, g: float = GRAVITY, limit: float = 1.5,
                    iterations: int = 10) -> np.ndarray:
    """theta_rest [W, P, 2].  Newton's method with a finite-difference Jacobian (4 P residual evaluations per step) and
    a FIXED number of steps, so that the result of a world does not depend on which other worlds share its batch.
    Converges to ~1e-14 rad for the reference's plants (sag <= 0.2 rad)."""
    W, P = origin_R.shape[0], len(parent)
    mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (P,))
    theta = np.zeros((W, P, 2))
    if g == 0.0 or P == 0:
        return theta
    eps = 1e-6
    J = None
    for it in range(iterations):
        F = static_residual(parent, origin_R, origin_t, com_z, theta, kp, mass, g)
        if True:
            J = np.zeros((W, 2 * P, 2 * P))
            for k in range(2 * P):
                d = np.zeros((1, P, 2)); d.reshape(-1)[k] = eps
                J[:, :, k] = (static_residual(parent, origin_R, origin_t, com_z, theta + d, kp, mass, g) -
                              static_residual(parent, origin_R, origin_t, com_z, theta - d, kp, mass, g)) / (2 * eps)
        step = np.linalg.solve(J, -F[:, :, None])[:, :, 0]
        step = np.clip(step, -0.3, 0.3)
        theta = np.clip(theta + step.reshape(W, P, 2), -limit, limit)
    # A plant heavier than its springs carry (the upright pose is past buckling: e.g. 20 stems of 10 kg on one 500 N m
    # root spring) has no small-deflection balance; Newton then ends at the limits or does not settle.  Such worlds
    # keep theta = 0 (decided per world, so still independent of the batch).
    F = static_residual(parent, origin_R, origin_t, com_z, theta, kp, mass, g)
    bad = (np.max(np.abs(F), axis=1) > REST_TOL * kp) | (np.max(np.abs(theta), axis=(1, 2)) > REST_MAX_ANGLE)
    theta[bad] = 0.0
    return theta


def rest_pose(world, kp: float = SPRING_KP, g: float = GRAVITY) -> np.ndarray:
    """theta_rest [P, 2] of one PlantWorld (all its plants: stems of different plants do not interact)."""
    P = world.num_stems
    if P == 0:
        return np.zeros((0, 2))
    parent = [s.parent for s in world.stems]
    oR = np.stack([s.origin.R for s in world.stems])[None]
    ot = np.stack([s.origin.t for s in world.stems])[None]
    cz = np.array([s.com_z for s in world.stems])[None]
    mass = np.array([getattr(s, "mass", STEM_MASS) for s in world.stems])
    lim = min(s.limit for s in world.stems)
    return rest_pose_batch(parent, oR, ot, cz, kp, mass, g, lim)[0]
