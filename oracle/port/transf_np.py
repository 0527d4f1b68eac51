"""TEST INFRASTRUCTURE ONLY -- oracle restatement (numpy, fp64, batched) of GraspIt!'s ``transf``.

Follows /root/reference/include/graspit/transform.h:12-93 and src/transform.cpp:
  * a transf is (unit quaternion, translation); every constructor re-normalises the quaternion
    (transform.h:47-49 ``set``),
  * ``a % b`` = (qa*qb, ta + qa*tb)                      (transform.cpp:154-159),
  * ``a * p`` = qa*p + ta with Eigen's q*v formula         (transform.cpp:164-168),
  * ``inverse`` = (q^-1, -(q^-1 * t))                      (transform.cpp:68-72),
  * ``AXIS_ANGLE_ROTATION`` returns IDENTITY when angle^2 < 1e-6 (transform.cpp:26-29).
All arrays carry a leading batch axis: q (N,4) as wxyz, t (N,3).
Pinned against the compiled reference (oracle/_ref) by tests/test_oracle_port.py.
"""
import numpy as np


def qnorm(q):
    n = np.sqrt(np.sum(q * q, axis=-1, keepdims=True))
    return np.where(n > 0, q / np.where(n > 0, n, 1.0), q)


def qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx], axis=-1)


def qrot(q, v):
    u = q[..., 1:4]
    uv = np.cross(u, v)
    uv = uv + uv
    return v + q[..., 0:1] * uv + np.cross(u, uv)


def qinv(q):
    n2 = np.sum(q * q, axis=-1, keepdims=True)
    return np.concatenate([q[..., 0:1], -q[..., 1:4]], axis=-1) / n2


def make(q, t):
    q = np.asarray(q, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    return qnorm(q), t


def compose(a, b):
    qa, ta = a
    qb, tb = b
    qa, qb = np.broadcast_arrays(qa, qb)
    return qnorm(qmul(qa, qb)), ta + qrot(qa, tb)


def inverse(a):
    q, t = a
    qi = qinv(q)
    return qnorm(qi), -qrot(qi, t)


def apply(a, p):
    return qrot(a[0], p) + a[1]


def rotate(a, v):
    """transf.affine() * v : rotation matrix of the (normalised) quaternion times v."""
    return qrot(a[0], v)


def identity(n=1):
    q = np.zeros((n, 4)); q[:, 0] = 1.0
    return q, np.zeros((n, 3))


def translation(v, n=None):
    v = np.atleast_2d(np.asarray(v, dtype=np.float64))
    q = np.zeros((v.shape[0], 4)); q[:, 0] = 1.0
    return q, v


def axis_angle_rotation(angle, axis):
    """Batched transf::AXIS_ANGLE_ROTATION; ``angle`` (N,), ``axis`` (3,) or (N,3)."""
    angle = np.atleast_1d(np.asarray(angle, dtype=np.float64))
    axis = np.asarray(axis, dtype=np.float64)
    axis = np.broadcast_to(axis, (angle.shape[0], 3))
    n = np.linalg.norm(axis, axis=-1, keepdims=True)
    axis = axis / np.where(n > 0, n, 1.0)
    h = 0.5 * angle
    q = np.concatenate([np.cos(h)[:, None], np.sin(h)[:, None] * axis], axis=-1)
    q = qnorm(q)
    snap = (angle * angle) < 0.000001
    q[snap] = np.array([1.0, 0.0, 0.0, 0.0])
    return q, np.zeros((angle.shape[0], 3))
