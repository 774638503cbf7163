"""Small float64 rigid-transform helpers used by the model compiler (host side only).

Conventions follow PyBullet, which is what the reference talks to:
  * quaternions are (x, y, z, w);
  * URDF ``rpy`` / ``getQuaternionFromEuler`` are fixed-axis roll-pitch-yaw, i.e.
    R = Rz(yaw) @ Ry(pitch) @ Rx(roll)  (reference use: rand_gen_plants_programmatically_main.py:137,165,215,245).
"""
from __future__ import annotations

import math

import numpy as np


def rot_x(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def rot_y(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def rot_z(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rpy_to_matrix(rpy) -> np.ndarray:
    r, p, y = (float(v) for v in rpy)
    return rot_z(y) @ rot_y(p) @ rot_x(r)


def axis_angle_matrix(axis, angle: float) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    k = skew(a)
    return np.eye(3) + math.sin(angle) * k + (1.0 - math.cos(angle)) * (k @ k)


def skew(v) -> np.ndarray:
    x, y, z = (float(c) for c in v)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def quat_from_rpy(rpy) -> np.ndarray:
    """(x, y, z, w) of Rz(yaw)Ry(pitch)Rx(roll) -- Bullet's setEulerZYX."""
    r, p, y = (float(v) * 0.5 for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
        cr * cp * cy + sr * sp * sy,
    ])


def quat_to_matrix(q) -> np.ndarray:
    x, y, z, w = (float(c) for c in q)
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([
        [1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
        [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
        [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)],
    ])


def matrix_to_quat(m) -> np.ndarray:
    """Rotation matrix -> (x, y, z, w), w >= 0 branch-stable (Shepperd)."""
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0.0:
        s = math.sqrt(t + 1.0) * 2.0
        q = np.array([(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2.0
        q = np.array([0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s])
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2.0
        q = np.array([(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s])
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2.0
        q = np.array([(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s])
    if q[3] < 0.0:
        q = -q
    return q


class Frame:
    """Rigid transform x -> R x + t."""

    __slots__ = ("R", "t")

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3)

    @staticmethod
    def from_xyz_rpy(xyz=(0.0, 0.0, 0.0), rpy=(0.0, 0.0, 0.0)) -> "Frame":
        return Frame(rpy_to_matrix(rpy), np.asarray(xyz, dtype=np.float64))

    def __matmul__(self, other: "Frame") -> "Frame":
        return Frame(self.R @ other.R, self.R @ other.t + self.t)

    def apply(self, p) -> np.ndarray:
        return self.R @ np.asarray(p, dtype=np.float64) + self.t

    def inverse(self) -> "Frame":
        return Frame(self.R.T, -self.R.T @ self.t)

    def copy(self) -> "Frame":
        return Frame(self.R.copy(), self.t.copy())
