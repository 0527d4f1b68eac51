"""Rigid transforms as (unit quaternion wxyz, translation), fp64, host side.

Same algebra as GraspIt!'s ``transf`` (reference ``include/graspit/transform.h:12-93``,
``src/transform.cpp``): ``compose(a, b)`` is ``a % b`` (apply ``b`` first, then ``a``);
``apply(a, p)`` is ``a * p``.  Used by the asset loaders to evaluate ``<transform>`` elements.
"""
from __future__ import annotations

import math

import numpy as np

IDENTITY = (np.array([1.0, 0.0, 0.0, 0.0]), np.zeros(3))


def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ])


def qnormalize(q):
    q = np.asarray(q, dtype=np.float64)
    n = math.sqrt(float(q @ q))
    return q / n if n > 0 else q


def qconj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def qrot(q, v):
    u = np.asarray(q[1:4], dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    uv = 2.0 * np.cross(u, v)
    return v + q[0] * uv + np.cross(u, uv)


def qmat(q):
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def mat2q(m):
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)  # w x y z
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def make(q, t):
    return (qnormalize(q), np.asarray(t, dtype=np.float64).copy())


def from_matrix(m, t):
    return (qnormalize(mat2q(m)), np.asarray(t, dtype=np.float64).copy())


def translation(v):
    return (np.array([1.0, 0.0, 0.0, 0.0]), np.asarray(v, dtype=np.float64).copy())


def angle_axis(angle, axis):
    """Plain Eigen::AngleAxisd -> quaternion (no small-angle snap)."""
    a = np.asarray(axis, dtype=np.float64)
    s = math.sin(0.5 * angle)
    return (qnormalize(np.array([math.cos(0.5 * angle), s * a[0], s * a[1], s * a[2]])), np.zeros(3))


def axis_angle_rotation(angle, axis):
    """transf::AXIS_ANGLE_ROTATION (src/transform.cpp:23-34): IDENTITY when angle^2 < 1e-6."""
    if angle * angle < 0.000001:
        return (np.array([1.0, 0.0, 0.0, 0.0]), np.zeros(3))
    a = np.asarray(axis, dtype=np.float64)
    n = math.sqrt(float(a @ a))
    if n > 0:
        a = a / n
    return angle_axis(angle, a)


def compose(a, b):
    """a % b (src/transform.cpp:154-159)."""
    return (qnormalize(qmul(a[0], b[0])), a[1] + qrot(a[0], b[1]))


def inverse(a):
    """src/transform.cpp:68-72."""
    qi = qconj(a[0]) / float(a[0] @ a[0])
    return (qnormalize(qi), -qrot(qi, a[1]))


def apply(a, p):
    return qrot(a[0], p) + a[1]


def to7(a):
    return np.concatenate([a[0], a[1]])
