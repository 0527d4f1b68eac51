"""Seeded synthetic pose batches of the benchmark configurations; numpy only, no geometry.

Used by bench.py, the tools and the tests alike, so that the product side of a measurement depends on nothing
outside this package."""
from __future__ import annotations

import numpy as np


def hand_robot(scene):
    """index of the robot that carries the virtual contacts (the hand of an arm + hand tree)"""
    return 0 if len(scene.robots) == 1 else max(range(len(scene.robots)), key=lambda i: len(scene.robots[i].vcontacts))


def object_body(scene):
    for i, b in enumerate(scene.bodies):
        if b.kind == "object":
            return i
    return -1


def object_centroid(scene, obj):
    return scene.bodies[obj].tris.reshape(-1, 3).astype(np.float64).mean(axis=0)


def sample_aa_states(scene, n, seed=1234, strata=("far", "near"), ridx=None):
    """SURVEY 8(d) config 1 sampling: SPACE_AXIS_ANGLE position variables relative to the object (planner box
    searchStateImpl.cpp:147-155 = "far"; a 150 mm ball about the object's centroid = "near"; 80 mm = "close") + raw DOFs
    uniform in their limits.  Returns (pos6 float32 (n,6), dofs float32 (n,ndof))."""
    rng = np.random.default_rng(seed)
    ridx = hand_robot(scene) if ridx is None else ridx
    r = scene.robots[ridx]
    obj = object_body(scene)
    c = object_centroid(scene, obj)
    per = [n // len(strata)] * len(strata)
    per[0] += n - sum(per)
    pos = []
    for s, k in zip(strata, per):
        th = rng.uniform(0, np.pi, k); ph = rng.uniform(-np.pi, np.pi, k); al = rng.uniform(0, np.pi, k)
        if s == "far":
            t = np.stack([rng.uniform(-250, 250, k), rng.uniform(-250, 250, k), rng.uniform(-250, 350, k)], 1)
        else:
            v = rng.normal(size=(k, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
            rad = (150.0 if s == "near" else 80.0) * rng.uniform(0, 1, k) ** (1.0 / 3.0)
            t = c[None, :] + v * rad[:, None]
        pos.append(np.concatenate([t, th[:, None], ph[:, None], al[:, None]], 1))
    pos = np.concatenate(pos, 0).astype(np.float32)
    dofs = rng.uniform(r.dof_min, r.dof_max, size=(n, r.n_dof)).astype(np.float32)
    return pos, dofs


def sample_planner_states(scene, n, seed, strata=("far", "near")):
    """(n, 6 + nEG) float32 planner states (SPACE_AXIS_ANGLE + POSE_EIGEN): Tx Ty Tz theta phi alpha + eigengrasp
    amplitudes uniform in [-4, 4] (PostureStateEigen::createVariables, searchStateImpl.cpp:50-80)."""
    pos, _ = sample_aa_states(scene, n, seed=seed, strata=strata)
    rng = np.random.default_rng(seed + 7919)
    neg = scene.robots[hand_robot(scene)].eg.shape[0]
    amps = rng.uniform(-4.0, 4.0, size=(n, neg)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([pos, amps], 1).astype(np.float32))
