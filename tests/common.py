"""Shared helpers of the parity tests: seeded pose batches, the oracle pipeline, the engine pipeline."""
from __future__ import annotations

import numpy as np

from graspit_b200.formats.scenepack import load_pack
from graspit_b200.sampling import hand_robot, object_body, object_centroid, sample_aa_states  # noqa: F401  (re-exported)
from oracle import ref
from oracle.port import fk as ofk
from oracle.port import transf_np as T


def aa_to_raw_states(scene, pos6, dofs, ridx=None):
    """(pos6, dofs) -> RAW engine input (n, 7+ndof) float32: base from the oracle's state mapping, rounded."""
    ridx = hand_robot(scene) if ridx is None else ridx
    r = scene.robots[ridx]
    obj = object_body(scene)
    base = ofk.aa_state_to_base(r, scene.body_tran[obj], pos6.astype(np.float64))
    raw = np.concatenate([base[0], base[1], dofs.astype(np.float64)], 1).astype(np.float32)
    return raw


class OracleScene:
    """The scene inside the reference backend (oracle/_ref), one world per thread."""

    def __init__(self, scene, nthreads=1, apply_touching=True):
        self.scene = scene
        self.worlds = []
        for _ in range(nthreads):
            w = ref.RefWorld()
            ids = {}
            for i, b in enumerate(scene.bodies):
                ids[i] = w.add_body(b.tris.astype(np.float64))
                w.set_transform(ids[i], *scene.body_tran[i])
            for (a, b) in scene.disabled_pairs:
                w.activate_pair(ids[a], ids[b], False)
            if apply_touching:
                for (a, b) in scene.touching_pairs:
                    w.activate_pair(ids[a], ids[b], False)
            self.worlds.append(w)
            self.ids = ids
        self.w = self.worlds[0]

    def link_poses(self, ridx, raw_states):
        """oracle FK on the widened fp32 RAW states -> (n, nbodies, 7) in robot.body_list() order (+children)"""
        sc = self.scene
        raw = np.asarray(raw_states, dtype=np.float64)
        base = T.make(raw[:, 0:4], raw[:, 4:7])
        order, dof_map, ofs = [], {}, 7
        stack = [ridx]
        robots = []
        while stack:
            k = stack.pop(0)
            robots.append(k)
            for c in sc.robots[k].chains:
                for (child, _) in c.children:
                    stack.append(child)
        for k in robots:
            nd = sc.robots[k].n_dof
            dof_map[k] = raw[:, ofs:ofs + nd]
            ofs += nd
            order += sc.robots[k].body_list()
        for k in range(len(sc.robots)):
            dof_map.setdefault(k, np.zeros((raw.shape[0], sc.robots[k].n_dof)))
        poses = ofk.forward_kinematics(sc.robots, ridx, base, dof_map)
        tf7 = np.stack([np.concatenate([poses[b][0], poses[b][1]], 1) for b in order], 1)
        return order, tf7

    def eval(self, ridx, obj, raw_states, mode=0, want_dist=False):
        sc = self.scene
        order, tf7 = self.link_poses(ridx, raw_states)
        r = sc.robots[ridx]
        hand_ids = [self.ids[b] for b in order]
        vc_link = [order.index(v.body) for v in r.vcontacts]
        vc_loc = np.array([v.loc for v in r.vcontacts]).reshape(-1, 3)
        vc_n = np.array([v.normal for v in r.vcontacts]).reshape(-1, 3)
        nroot = len(r.body_list())
        if mode == 1 or want_dist:
            # LIVE contacts / link distances are defined over the hand's own bodies only
            assert nroot == len(order), "LIVE oracle supports single-robot trees"
        legal, energy, dist = ref.eval_batch(self.worlds, hand_ids, self.ids[obj], tf7, vc_link, vc_loc, vc_n,
                                             mode=mode, want_dist=want_dist)
        return legal, energy, dist, order, tf7

    def pair_sets(self, ridx, raw_states):
        """per pose: set of colliding (scene body a, scene body b) among pairs touching the robot"""
        order, tf7 = self.link_poses(ridx, raw_states)
        inv = {v: k for k, v in self.ids.items()}
        interest = [self.ids[b] for b in order]
        out = []
        for p in range(tf7.shape[0]):
            for k, b in enumerate(order):
                self.w.set_transform(self.ids[b], tf7[p, k, 0:4], tf7[p, k, 4:7])
            _, pairs = self.w.all_collisions(fast=False, interest=interest)
            out.append({tuple(sorted((inv[a], inv[b]))) for a, b in pairs})
        return out
