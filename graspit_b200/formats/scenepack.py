"""Scene packs: a whole GraspIt! world (triangle soups, robot descriptions, pair masks) in one ``.npz``.

GraspIt!'s asset tree (``models/``, ``worlds/``) is not available where the engine is benchmarked, so
``tools/make_scenes.py`` reads the XML / Inventor / VRML files once (``formats.xmlmodel``) and stores the
result in this compact interchange format; ``load_pack`` restores the same ``Scene`` dataclasses.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .xmlmodel import (BodyDesc, ChainDesc, JointDesc, RobotDesc, Scene, VirtualContactDesc)

SCENE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scenes")


def _tf(t):
    return [list(map(float, t[0])), list(map(float, t[1]))]


def _untf(t):
    return (np.array(t[0], dtype=np.float64), np.array(t[1], dtype=np.float64))


def save_pack(scene: Scene, path: str, extra: dict | None = None):
    arrays = {}
    meta = {"name": scene.name, "bodies": [], "robots": [], "body_tran": [_tf(t) for t in scene.body_tran],
            "disabled_pairs": [list(p) for p in scene.disabled_pairs],
            "touching_pairs": [list(p) for p in scene.touching_pairs], "extra": extra or {}}
    # identical meshes (e.g. the three Barrett link2 copies) are stored once
    seen = {}
    for i, b in enumerate(scene.bodies):
        key = (b.tris.shape, b.tris.tobytes())
        if key not in seen:
            seen[key] = f"tris_{len(seen)}"
            arrays[seen[key]] = b.tris.astype(np.float32)
        meta["bodies"].append({"name": b.name, "kind": b.kind, "robot": b.robot, "chain": b.chain, "link": b.link,
                               "youngs": b.youngs, "material": b.material, "tris": seen[key]})
    for r in scene.robots:
        meta["robots"].append({
            "name": r.name, "base_body": r.base_body, "mount_body": r.mount_body, "n_dof": r.n_dof,
            "dof_type": r.dof_type, "dof_min": list(map(float, r.dof_min)), "dof_max": list(map(float, r.dof_max)),
            "dof_default": list(map(float, r.dof_default)), "dof_values": list(map(float, r.dof_values)),
            "dof_velocity": None if r.dof_velocity is None else list(map(float, r.dof_velocity)),
            "tran": _tf(r.tran), "approach": _tf(r.approach), "parent": r.parent, "parent_chain": r.parent_chain,
            "eg": None if r.eg is None else [list(map(float, e)) for e in r.eg],
            "eg_origin": None if r.eg_origin is None else list(map(float, r.eg_origin)),
            "eg_norm": None if r.eg_norm is None else list(map(float, r.eg_norm)),
            "eg_limits": None if r.eg_limits is None else [[None if np.isnan(v) else float(v) for v in e] for e in r.eg_limits],
            "chains": [{"tran": _tf(c.tran), "link_bodies": c.link_bodies, "last_joint": c.last_joint,
                        "children": [[ci, _tf(o)] for ci, o in c.children],
                        "joints": [[j.dof, j.ratio, j.offset, bool(j.revolute), j.theta, j.d, j.a, j.alpha, j.min, j.max]
                                   for j in c.joints]} for c in r.chains],
            "vcontacts": [[v.body, v.finger, v.link, list(map(float, v.loc)), list(map(float, v.normal)), _tf(v.frame)]
                          for v in r.vcontacts],
        })
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **arrays)


def subdivide(tris: np.ndarray, levels: int) -> np.ndarray:
    """midpoint subdivision in fp32, deterministic (config 5 object: mug x 4^3 = 220 800 triangles)"""
    t = np.asarray(tris, dtype=np.float32)
    half = np.float32(0.5)
    for _ in range(levels):
        a, b, c = t[:, 0], t[:, 1], t[:, 2]
        ab = ((a + b) * half).astype(np.float32)
        bc = ((b + c) * half).astype(np.float32)
        ca = ((c + a) * half).astype(np.float32)
        t = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1),
                            np.stack([ab, bc, ca], 1)], 0)
    return np.ascontiguousarray(t)


def load_pack(path_or_name: str) -> Scene:
    if path_or_name == "barrett_mug_200k":
        # the plannerMug world with the mug midpoint-subdivided 3x; derived, not stored
        sc = load_pack("barrett_mug")
        sc.name = "barrett_mug_200k"
        sc.bodies[0].tris = subdivide(sc.bodies[0].tris, 3)
        sc.bodies[0].name = "mug_x64"
        return sc
    path = path_or_name
    if not os.path.exists(path):
        path = os.path.join(SCENE_DIR, path_or_name if path_or_name.endswith(".npz") else path_or_name + ".npz")
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    sc = Scene(name=meta["name"])
    for b in meta["bodies"]:
        sc.bodies.append(BodyDesc(name=b["name"], tris=np.ascontiguousarray(z[b["tris"]], dtype=np.float32), kind=b["kind"],
                                  robot=b["robot"], chain=b["chain"], link=b["link"], youngs=b["youngs"],
                                  material=b["material"]))
    sc.body_tran = [_untf(t) for t in meta["body_tran"]]
    sc.disabled_pairs = [tuple(p) for p in meta["disabled_pairs"]]
    sc.touching_pairs = [tuple(p) for p in meta["touching_pairs"]]
    for r in meta["robots"]:
        chains = []
        for c in r["chains"]:
            joints = [JointDesc(int(j[0]), j[1], j[2], bool(j[3]), j[4], j[5], j[6], j[7], j[8], j[9]) for j in c["joints"]]
            chains.append(ChainDesc(_untf(c["tran"]), joints, list(c["link_bodies"]), list(c["last_joint"]),
                                    [(ci, _untf(o)) for ci, o in c["children"]]))
        rd = RobotDesc(name=r["name"], base_body=r["base_body"], chains=chains, n_dof=r["n_dof"], dof_type=r["dof_type"],
                       dof_min=np.array(r["dof_min"]), dof_max=np.array(r["dof_max"]),
                       dof_default=np.array(r["dof_default"]), dof_values=np.array(r["dof_values"]),
                       tran=_untf(r["tran"]), approach=_untf(r["approach"]),
                       eg=None if r["eg"] is None else np.array(r["eg"]),
                       eg_origin=None if r["eg_origin"] is None else np.array(r["eg_origin"]),
                       eg_norm=None if r["eg_norm"] is None else np.array(r["eg_norm"]),
                       eg_limits=None if r["eg_limits"] is None else np.array(
                           [[np.nan if v is None else v for v in e] for e in r["eg_limits"]]),
                       mount_body=r["mount_body"], parent=r["parent"], parent_chain=r["parent_chain"],
                       dof_velocity=None if r.get("dof_velocity") is None else np.array(r["dof_velocity"]))
        for v in r["vcontacts"]:
            rd.vcontacts.append(VirtualContactDesc(v[0], v[1], v[2], np.array(v[3]), np.array(v[4]), _untf(v[5])))
        sc.robots.append(rd)
    sc.extra = meta.get("extra", {})
    return sc
