"""Time-stepped restatement of the reference's plant response.                 *** TEST INFRASTRUCTURE ***

Only tests/ and tools/ import this module.  It exists to MEASURE how far the stateless quasi-static model of
oracle/edgecheck_oracle.py (and of the CUDA kernels) is from the reference's dynamic procedure; it is independent of
that model (own kinematics, own Jacobians, own contact solve) and shares only the plant description (PlantWorld, whose
sampler is pinned bit for bit to the reference) and the arm's sphere centres.

What it restates (paths relative to /root/reference/plant_motion_planning/):
  env.py:146-163, 166-172, 212-227   step / step_after_restoring: optional teleport of the arm, then a position motor
                                     with gain 0.01, then 50 x (spring torques; stepSimulation) at 1/240 s
  env.py:127-144                     spring: EXTERNAL LINK TORQUE [-kp tx + kd (tx - tx_prev), -kp ty + kd (ty - ty_prev), 0]
                                     in the stem link's frame, kp = 500, kd = 50 (per call, not per second).  Being an
                                     external torque on the link, it also loads every ancestor joint (axis_j . torque)
  rand_gen_plants_programmatically_main.py:255-272, 296-307   stem mass 10 kg, inertial frame 0.25 m up the stem,
                                     massless v1 link, revolute x then y at the stem root, limits +-1.5, joint damping 10
  pybullet_tools/utils.py:1218-1226  gravity -9.8, default time step 1/240
The integrator lives in PyBullet 3.2.0 (Bullet btMultiBodyDynamicsWorld; not under /root/reference).  Restated from
Bullet's public algorithm descriptions, NOT verified against the engine (it cannot run here): generalized coordinates
(two angles per stem), M(q) qdd = Q with Q = gravity + external link torques + joint damping (velocity-product terms
are dropped: joint speeds stay below 1 rad/s), semi-implicit Euler (velocities first, contacts solved on the
velocities, then positions), contacts as a projected Gauss-Seidel loop on velocity constraints with Baumgarte
stabilisation erp = 0.2 (Bullet's m_erp2), the arm a kinematic body that closes 1 % of its joint error per sub-step
(btMultiBodyJointMotor: target velocity = kp * error / dt with erp 1), joint limits by clamping.  Friction optional
(default 0; Bullet's combined coefficient would be 0.5 * 0.5).  Box inertia from the stem's extents about its COM.
PARITY STATUS: unpinned (no artefact of the reference carries a deflection value); the numbers produced with this
module are an estimate of the modelling gap, stated as such in DESIGN.md.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class DynParams:
    kp: float = 500.0            # env.py:128
    kd: float = 50.0             # env.py:129
    joint_damping: float = 10.0  # rand_gen...:307
    mass: float = 10.0           # rand_gen...:255
    g: float = 9.8               # pybullet_tools/utils.py:1218-1221
    dt: float = 1.0 / 240.0      # pybullet_tools/utils.py:56
    substeps: int = 50           # env.py:157, 216
    arm_gain: float = 0.01       # env.py:154, 170 (positionGains)
    erp: float = 0.2
    pgs_iterations: int = 30
    friction: float = 0.0
    gravity: bool = True
    ancestor_loading: bool = True   # False: the spring acts on the stem's own two joints only (an ideal joint spring)
    arm_lag: bool = True            # False: the arm is at the commanded configuration during the whole step
    contact_margin: float = 0.02    # contacts are generated inside Bullet's contact breaking threshold


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


class DynamicPlantWorld:
    """One PlantWorld (all its stems) as a tree of universal joints, pushed by kinematic spheres."""

    def __init__(self, world, sphere_radii: np.ndarray, prm: Optional[DynParams] = None):
        self.world = world
        self.prm = prm or DynParams()
        self.P = world.num_stems
        self.radii = np.asarray(sphere_radii, dtype=np.float64)
        self.theta = np.zeros((self.P, 2))
        self.omega = np.zeros((self.P, 2))
        self.prev = np.zeros((self.P, 2))        # env.py:118-119 prev_plant_rot_joint_displacement_{x,y}
        # chain of ancestors (self included), root first
        self.chain: List[List[int]] = []
        for s, st in enumerate(world.stems):
            self.chain.append(([] if st.parent < 0 else list(self.chain[st.parent])) + [s])
        # box inertia about the COM (extents 2w x 2w x 2h), link frame
        self.Ic = []
        for st in world.stems:
            w2, h2 = (2 * st.half_width) ** 2, (2 * st.half_height) ** 2
            m = self.prm.mass
            self.Ic.append(np.diag([m / 12 * (w2 + h2), m / 12 * (w2 + h2), m / 12 * (2 * w2)]))

    # ------------------------------------------------------------------ kinematics
    def kinematics(self, theta: np.ndarray):
        """Link frames (R, t), joint axes (world) and joint positions of every stem."""
        P = self.P
        R = [None] * P; t = [None] * P; ax = [None] * P; ay = [None] * P
        for s, st in enumerate(self.world.stems):
            if st.parent < 0:
                Rp, tp = np.eye(3), np.zeros(3)
            else:
                Rp, tp = R[st.parent], t[st.parent]
            Rv = Rp @ st.origin.R
            t[s] = Rp @ st.origin.t + tp
            Rx = Rv @ _rx(theta[s, 0])
            ax[s] = Rv[:, 0].copy()
            ay[s] = Rx[:, 1].copy()
            R[s] = Rx @ _ry(theta[s, 1])
        return R, t, ax, ay

    def _jac_point(self, s: int, p: np.ndarray, t, ax, ay) -> np.ndarray:
        """3 x 2P linear Jacobian of world point p rigidly attached to stem s."""
        J = np.zeros((3, 2 * self.P))
        for a in self.chain[s]:
            r = p - t[a]
            J[:, 2 * a] = np.cross(ax[a], r)
            J[:, 2 * a + 1] = np.cross(ay[a], r)
        return J

    def _jac_rot(self, s: int, ax, ay) -> np.ndarray:
        J = np.zeros((3, 2 * self.P))
        for a in self.chain[s]:
            J[:, 2 * a] = ax[a]
            J[:, 2 * a + 1] = ay[a]
        return J

    def mass_and_forces(self, theta, omega):
        prm = self.prm
        R, t, ax, ay = self.kinematics(theta)
        n = 2 * self.P
        M = np.zeros((n, n))
        Q = np.zeros(n)
        for s, st in enumerate(self.world.stems):
            com = R[s] @ np.array([0.0, 0.0, st.com_z]) + t[s]
            Jv = self._jac_point(s, com, t, ax, ay)
            Jw = self._jac_rot(s, ax, ay)
            Iw = R[s] @ self.Ic[s] @ R[s].T
            M += prm.mass * Jv.T @ Jv + Jw.T @ Iw @ Jw
            if prm.gravity:
                Q += Jv.T @ np.array([0.0, 0.0, -prm.mass * prm.g])
            # spring: external torque in the LINK frame (env.py:140-143)
            tl = np.array([-prm.kp * theta[s, 0] + prm.kd * (theta[s, 0] - self.prev[s, 0]),
                           -prm.kp * theta[s, 1] + prm.kd * (theta[s, 1] - self.prev[s, 1]), 0.0])
            tw = R[s] @ tl
            if prm.ancestor_loading:
                Q += Jw.T @ tw
            else:
                Q[2 * s] += ax[s] @ tw
                Q[2 * s + 1] += ay[s] @ tw
        Q -= prm.joint_damping * omega.reshape(-1)
        return M, Q, (R, t, ax, ay)

    # ------------------------------------------------------------------ contacts
    def _contacts(self, kin, centers: np.ndarray):
        """(stem, contact point on the box [world], normal box->sphere, gap) for sphere/box pairs inside the margin."""
        R, t, ax, ay = kin
        out = []
        for s, st in enumerate(self.world.stems):
            half = np.array([st.half_width, st.half_width, st.half_height])
            cb = np.array([0.0, 0.0, st.z_offset + st.half_height])
            pl = (centers - t[s][None]) @ R[s] - cb[None]          # sphere centres in the box frame
            q = np.clip(pl, -half, half)
            d = pl - q
            dist = np.linalg.norm(d, axis=1)
            gap = dist - self.radii
            for a in np.nonzero(gap < self.prm.contact_margin)[0]:
                if dist[a] < 1e-9:
                    continue                                        # centre inside the box: no usable normal
                nl = d[a] / dist[a]
                out.append((s, R[s] @ (q[a] + cb) + t[s], R[s] @ nl, float(gap[a]), int(a)))
        return out

    # ------------------------------------------------------------------ one sub-step
    def substep(self, centers: np.ndarray, centers_next: np.ndarray):
        prm = self.prm
        dt = prm.dt
        M, Q, kin = self.mass_and_forces(self.theta, self.omega)
        self.prev = self.theta.copy()            # env.py:137-138: updated on every _restore_plant_joints call
        Minv = np.linalg.inv(M + 1e-9 * np.eye(M.shape[0]))
        w = self.omega.reshape(-1) + dt * (Minv @ Q)
        cons = self._contacts(kin, centers)
        if cons:
            R, t, ax, ay = kin
            vs = (centers_next - centers) / dt
            rows = []
            for (s, pc, n, gap, a) in cons:
                Jn = n @ self._jac_point(s, pc, t, ax, ay)          # stem point velocity along n
                target = -gap / dt if gap > 0 else prm.erp * (-gap) / dt
                # gap rate = n.v_sphere - Jn.w  >= target
                meff = 1.0 / max(float(Jn @ Minv @ Jn), 1e-12)
                rows.append([Jn, float(n @ vs[a]), target, meff, 0.0])
            for _ in range(prm.pgs_iterations):
                for row in rows:
                    Jn, vsn, target, meff, lam = row
                    rate = vsn - float(Jn @ w)
                    dl = (target - rate) * meff
                    new = max(0.0, lam + dl)
                    # impulse -n * lam on the stem: w += Minv Jn^T (-(new - lam))
                    w -= (Minv @ Jn) * (new - lam)
                    row[4] = new
        self.omega = w.reshape(self.P, 2)
        th = self.theta + dt * self.omega
        for s, st in enumerate(self.world.stems):
            for k in range(2):
                if abs(th[s, k]) > st.limit:
                    th[s, k] = np.sign(th[s, k]) * st.limit
                    self.omega[s, k] = 0.0
        self.theta = th

    def settle(self, centers: np.ndarray, substeps: int = 600):
        for _ in range(substeps):
            self.substep(centers, centers)

    def frames(self):
        R, t, _, _ = self.kinematics(self.theta)
        return np.stack(R)[None], np.stack(t)[None]


class DynamicEnv:
    """MultiPlantWorld.step / step_after_restoring for a list of worlds sharing one arm (env.py:187-227)."""

    def __init__(self, model, worlds: Sequence, prm: Optional[DynParams] = None):
        from oracle import edgecheck_oracle as O
        self._O = O
        self.model = model
        self.prm = prm or DynParams()
        self.worlds = [DynamicPlantWorld(w, model.sph_radius, self.prm) for w in worlds]
        self.arm = np.zeros(model.dof)

    def centers(self, q) -> np.ndarray:
        return self._O.sphere_centers(self.model, np.asarray(q, dtype=np.float64)[None])[0]

    def reset(self, q, settle_substeps: int = 600):
        """Plants at theta = 0 (where the observers capture their reference pose, representation.py:206-212), arm
        teleported to q, then left to settle under gravity / springs / contact."""
        self.arm = np.asarray(q, dtype=np.float64).copy()
        c = self.centers(self.arm)
        for w in self.worlds:
            w.theta[:] = 0; w.omega[:] = 0; w.prev[:] = 0
            w.settle(c, settle_substeps)

    def step(self, q, set_joint_pos: bool = False):
        """env.py:212-219: 50 sub-steps towards the commanded configuration q."""
        q = np.asarray(q, dtype=np.float64)
        if set_joint_pos or not self.prm.arm_lag:
            self.arm = q.copy()
        c = self.centers(self.arm)
        for _ in range(self.prm.substeps):
            nxt = self.arm + self.prm.arm_gain * (q - self.arm) if self.prm.arm_lag else self.arm
            cn = self.centers(nxt)
            for w in self.worlds:
                w.substep(c, cn)
            self.arm, c = nxt, cn

    def deflections(self) -> List[np.ndarray]:
        """Chained observer values per world (representation.py:268-294 through the oracle's pinned observers)."""
        out = []
        for w in self.worlds:
            Rs, ts = w.frames()
            raw = self._O.raw_deflections(w.world, Rs, ts)
            out.append(self._O.chain_deflections(w.world, raw)[0])
        return out

    def thetas(self) -> List[np.ndarray]:
        return [w.theta.copy() for w in self.worlds]


def static_rest_pose(world, prm: Optional[DynParams] = None, tol: float = 1e-11) -> np.ndarray:
    """The pose in which springs and gravity balance (no contact): root of this module's own generalized force by
    Newton with a finite-difference Jacobian.  Independent of plant_motion_planning_b200/model/statics.py (different
    kinematics / Jacobian code); the two must agree (tests/test_dynamic_gap.py)."""
    prm = prm or DynParams()
    w = DynamicPlantWorld(world, np.zeros(1), DynParams(**{**prm.__dict__, "kd": 0.0, "joint_damping": 0.0}))
    P = w.P
    th = np.zeros((P, 2))
    if P == 0 or not prm.gravity:
        return th

    def force(t):
        w.prev = t.copy()
        return w.mass_and_forces(t, np.zeros((P, 2)))[1]

    eps = 1e-6
    for _ in range(50):
        F = force(th)
        if np.max(np.abs(F)) < tol * prm.kp:
            break
        J = np.zeros((2 * P, 2 * P))
        for k in range(2 * P):
            d = np.zeros(2 * P); d[k] = eps
            J[:, k] = (force(th + d.reshape(P, 2)) - force(th - d.reshape(P, 2))) / (2 * eps)
        step = np.clip(np.linalg.solve(J, -F), -0.3, 0.3)
        th = th + step.reshape(P, 2)
    # no small-deflection balance (plant past buckling): the model keeps theta = 0 (same rule as model/statics.py)
    if np.max(np.abs(force(th))) > 1e-7 * prm.kp or np.max(np.abs(th)) > 0.75:
        return np.zeros((P, 2))
    return th
