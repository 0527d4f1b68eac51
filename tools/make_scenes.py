#!/usr/bin/env python
"""Builds the scene packs under graspit_b200/scenes/ from the GraspIt! asset tree.

Run where /root/reference exists:  python tools/make_scenes.py [--graspit /root/reference]
For each world it (1) parses the world / robot / body XML and the Inventor / VRML geometry
(formats.xmlmodel), (2) reproduces World::addRobot's data-dependent pair disabling
(reference src/world.cpp:1080-1086: every pair found in contact when a robot is added is disabled)
with the oracle -- the reference's own allContacts -- and (3) writes the .npz pack.
tests/test_scene_packs.py re-derives the touching pairs with the CUDA engine and compares.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graspit_b200.formats import transf as tf                      # noqa: E402
from graspit_b200.formats.scenepack import SCENE_DIR, save_pack    # noqa: E402
from graspit_b200.formats.xmlmodel import BodyDesc, Scene, load_body, load_robot, load_world  # noqa: E402
from oracle import ref                                              # noqa: E402
from oracle.port import fk as ofk                                   # noqa: E402
from oracle.port import transf_np as T                              # noqa: E402

CONTACT_THRESHOLD = 0.2      # Contact::THRESHOLD, src/contact/contact.cpp:58


def _robot_poses(scene, ridx, base, dofs):
    """link poses of robot ridx (+ attached robots) for a single configuration"""
    dof_map = {i: np.asarray(scene.robots[i].dof_values if i != ridx else dofs, dtype=np.float64)[None, :]
               for i in range(len(scene.robots))}
    b = (np.asarray(base[0], dtype=np.float64)[None, :], np.asarray(base[1], dtype=np.float64)[None, :])
    out = ofk.forward_kinematics(scene.robots, ridx, b, dof_map)
    return {k: (v[0][0], v[1][0]) for k, v in out.items()}


def settle_scene(scene: Scene):
    """Fill body_tran of robot bodies (load poses) and touching_pairs, emulating the order of
    World::loadFromXml: a robot is added at identity with default DOF values (importRobot -> addRobot),
    the contact scan runs, and only then its <dofValues>/<transform> are applied."""
    w = ref.RefWorld()
    ids = {}
    added = []

    def add(i):
        ids[i] = w.add_body(scene.bodies[i].tris.astype(np.float64))
        q, t = scene.body_tran[i]
        w.set_transform(ids[i], q, t)
        added.append(i)

    def disabled(a, b):
        return (min(a, b), max(a, b)) in set(scene.disabled_pairs) | set(scene.touching_pairs)

    robots_of_body = {}
    for ri, r in enumerate(scene.robots):
        for b in r.body_list():
            robots_of_body[b] = ri
    order = list(range(len(scene.bodies)))
    done_robots = set()
    for i in order:
        if i in ids:
            continue
        ri = robots_of_body.get(i, -1)
        if ri < 0 or scene.bodies[i].kind == "mount":
            add(i)
            continue
        if ri in done_robots:
            continue
        r = scene.robots[ri]
        own = [b for b in r.body_list() if scene.bodies[b].kind != "mount"]
        # at addRobot time: identity base, default DOFs
        poses = _robot_poses(scene, ri, tf.IDENTITY, r.dof_default)
        for b in own:
            scene.body_tran[b] = poses[b]
            add(b)
        for (a, b) in scene.disabled_pairs:
            if a in ids and b in ids:
                w.activate_pair(ids[a], ids[b], False)
        for (a, b) in scene.touching_pairs:
            if a in ids and b in ids:
                w.activate_pair(ids[a], ids[b], False)
        inv = {v: k for k, v in ids.items()}
        for (pa, pb), contacts in w.all_contacts(CONTACT_THRESHOLD):
            p = (min(inv[pa], inv[pb]), max(inv[pa], inv[pb]))
            if p not in scene.touching_pairs:
                scene.touching_pairs.append(p)
            w.activate_pair(pa, pb, False)
        # now the world file's pose
        if r.parent < 0:
            poses = _robot_poses(scene, ri, r.tran, r.dof_values)
            for b, p in poses.items():
                scene.body_tran[b] = p
                if b in ids:
                    w.set_transform(ids[b], p[0], p[1])
        done_robots.add(ri)
    # attached robots follow their parents
    for ri, r in enumerate(scene.robots):
        if r.parent < 0:
            poses = _robot_poses(scene, ri, r.tran, r.dof_values)
            for b, p in poses.items():
                scene.body_tran[b] = p
    w.close()
    return scene


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graspit", default="/root/reference")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    G = args.graspit
    os.makedirs(SCENE_DIR, exist_ok=True)
    todo = args.only.split(",") if args.only else ["barrett_mug", "humanhand_flask", "dlr_flask", "staubli_barrett"]

    if "barrett_mug" in todo:
        sc = settle_scene(load_world(os.path.join(G, "worlds/plannerMug.xml"), G))
        sc.name = "barrett_mug"
        save_pack(sc, os.path.join(SCENE_DIR, "barrett_mug.npz"), {"world": "worlds/plannerMug.xml"})
        print("barrett_mug:", len(sc.bodies), "bodies", sum(len(b.tris) for b in sc.bodies), "tris; touching", sc.touching_pairs)

    if "humanhand_flask" in todo:
        # config 3: flask at its dlr_flask.xml pose + HumanHand20DOF (built programmatically)
        sc = Scene(name="humanhand_flask")
        flask = load_body(os.path.join(G, "models/objects/flask.xml"), kind="object")
        sc.add_body(flask, tf.make([0.702043, -0.712135, 0, 0], [-124.849, 145.862, -0.000865237]))
        load_robot(sc, os.path.join(G, "models/robots/HumanHand/HumanHand20DOF.xml"), tf.IDENTITY)
        settle_scene(sc)
        save_pack(sc, os.path.join(SCENE_DIR, "humanhand_flask.npz"), {"world": "programmatic: flask + HumanHand20DOF"})
        print("humanhand_flask:", len(sc.bodies), "bodies", sum(len(b.tris) for b in sc.bodies), "tris; touching", sc.touching_pairs)

    if "dlr_flask" in todo:
        sc = settle_scene(load_world(os.path.join(G, "worlds/dlr_flask.xml"), G))
        sc.name = "dlr_flask"
        save_pack(sc, os.path.join(SCENE_DIR, "dlr_flask.npz"), {"world": "worlds/dlr_flask.xml"})
        print("dlr_flask:", len(sc.bodies), "bodies", sum(len(b.tris) for b in sc.bodies), "tris; touching", sc.touching_pairs)

    if "staubli_barrett" in todo:
        # config 4: arm + hand + obstacles at fixed poses (committed here)
        sc = Scene(name="staubli_barrett")
        obstacles = [("models/obstacles/floor.xml", tf.make([1, 0, 0, 0], [-2500.0, -4000.0, -5.0])),
                     ("models/obstacles/dinner_table.xml", tf.make([1, 0, 0, 0], [400.0, -485.0, -420.0])),
                     ("models/obstacles/shelf.xml", tf.make([0.70710678, 0, 0, 0.70710678], [-700.0, -400.0, -5.0]))]
        for fn, tr in obstacles:
            sc.add_body(load_body(os.path.join(G, fn), kind="obstacle"), tr)
        w = load_world(os.path.join(G, "worlds/staubli_barrett.xml"), G)
        off = len(sc.bodies)
        # append the world's bodies/robots after the obstacles, re-indexing body references
        for b, t in zip(w.bodies, w.body_tran):
            sc.add_body(b, t)
        for r in w.robots:
            r.base_body += off
            if r.mount_body >= 0:
                r.mount_body += off
            for c in r.chains:
                c.link_bodies = [x + off for x in c.link_bodies]
            for v in r.vcontacts:
                v.body += off
            sc.robots.append(r)
        sc.disabled_pairs = [(a + off, b + off) for a, b in w.disabled_pairs]
        settle_scene(sc)
        save_pack(sc, os.path.join(SCENE_DIR, "staubli_barrett.npz"),
                  {"world": "worlds/staubli_barrett.xml + obstacles", "obstacles": [o[0] for o in obstacles]})
        print("staubli_barrett:", len(sc.bodies), "bodies", sum(len(b.tris) for b in sc.bodies), "tris; touching", sc.touching_pairs)
    # barrett_mug_200k (config 5) is derived at load time: scenepack.load_pack('barrett_mug_200k')


if __name__ == "__main__":
    main()
