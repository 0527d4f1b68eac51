"""The product's replay of the reference's contact-candidate order (b2c_contact_visit_order, csrc/contacts_host.cpp)
against the reference itself: ContactCallback's own leaf-pair visit order, traced inside the verbatim-compiled backend
(oracle/ref_driver.cpp ref_contact_trace; reference src/Collision/Graspit/collisionAlgorithms.cpp:42-130).
Host-only: no GPU needed.  Poses = the closed hands of the autograsp golden fixture (every finger rests on the mug)."""
import os

import numpy as np
import pytest

from graspit_b200 import engine
from graspit_b200.formats.scenepack import load_pack
from oracle import ref
from oracle.port.autograsp import AutoGraspOracle
from tests.common import OracleScene, object_body

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "barrett_mug_autograsp.npz")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle library not built")


def _closed_hands(n):
    z = np.load(GOLD)
    ok = (~z["error"]) & z["moved"]
    return z["raw"][ok][:n].astype(np.float64), z["joints"][ok][:n]


def test_visit_order_matches_reference_trace():
    sc = load_pack("barrett_mug")
    osc = OracleScene(sc, nthreads=1)
    ag = AutoGraspOracle(sc, 0, osc.w, osc.ids)
    obj = object_body(sc)
    raw, joints = _closed_hands(24)
    hand = sc.robots[0].body_list()
    pairs = [(a, obj) for a in hand] + [(hand[i], hand[j]) for i in range(len(hand)) for j in range(i + 1, len(hand))]
    rng = np.random.default_rng(5)
    groups = multi = 0
    for p in range(len(raw)):
        ag.set_joints((raw[p, 0:4], raw[p, 4:7]), joints[p])
        tf = dict(ag.last_tf)
        tf[obj] = sc.body_tran[obj]
        for a, b in pairs:
            if not osc.w.is_active(osc.ids[a], osc.ids[b]):
                continue
            ta, tb = sc.bodies[a].tris.astype(np.float64), sc.bodies[b].tris.astype(np.float64)
            vis, rawc = osc.w.contact_trace(osc.ids[a], osc.ids[b], 0.2, ta, tb)
            if len(vis) == 0:
                continue
            groups += 1
            assert (vis[:, :2] >= 0).all()
            assert vis[:, 2].sum() == len(rawc)
            if len(vis) < 2:
                continue
            multi += 1
            perm = rng.permutation(len(vis))          # the device emits candidates in no particular order
            order = engine.contact_visit_order(sc.bodies[a].tris, sc.bodies[b].tris, tf[a], tf[b], vis[perm, 0], vis[perm, 1])
            assert (perm[order] == np.arange(len(vis))).all(), (p, a, b)
    assert groups >= 30 and multi >= 20, (groups, multi)


def test_visit_order_keeps_input_order_inside_a_triangle_pair():
    sc = load_pack("barrett_mug")
    obj = object_body(sc)
    a = sc.robots[0].body_list()[0]
    ta = np.array([3, 3, 3, 7, 7], dtype=np.int32)
    tb = np.array([10, 10, 10, 10, 10], dtype=np.int32)
    ident = (np.array([1.0, 0, 0, 0]), np.zeros(3))
    order = engine.contact_visit_order(sc.bodies[a].tris, sc.bodies[obj].tris, ident, ident, ta, tb)
    pos = {int(i): k for k, i in enumerate(order)}
    assert pos[0] < pos[1] < pos[2] and pos[3] < pos[4]
    assert sorted(order.tolist()) == [0, 1, 2, 3, 4]
