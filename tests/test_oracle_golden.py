"""The oracle (reference GraspitCollision compiled verbatim, oracle/_ref) against every golden value the
reference's own tests hold for this path: test/graspit_collision_tests.cpp,
test/transform_tests.cpp, test/eigen_tests.cpp.  CPU only."""
import math

import numpy as np
import pytest

from graspit_b200.formats.scenepack import load_pack
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
SLOP = 1e-3   # test/graspit_collision_tests.cpp:24


@pytest.fixture(scope="module")
def world():
    sc = load_pack("barrett_mug")
    w = ref.RefWorld()
    mug = w.add_body(sc.bodies[0].tris.astype(np.float64))
    palm = w.add_body(sc.bodies[1].tris.astype(np.float64))
    return sc, w, mug, palm


def _moved_mug(w, mug):
    # transf::AXIS_ANGLE_ROTATION(10, x) then TRANSLATION(1,2,3) % rot   (graspit_collision_tests.cpp:34-39)
    rot = ref.transf_axis_angle(10.0, [1, 0, 0])
    tr = ref.transf_compose(([1, 0, 0, 0], [1, 2, 3]), rot)
    w.set_transform(mug, tr[0], tr[1])


def test_body_to_body_distance_golden(world):
    sc, w, mug, palm = world
    _moved_mug(w, mug)
    w.set_transform(palm, [1, 0, 0, 0], [0, 0, 0])
    d, p1, p2 = w.body_distance(mug, palm)
    assert d == pytest.approx(-1, abs=SLOP)
    assert p1 == pytest.approx([-1, 3.31021, 1.42917], abs=SLOP)
    assert p2 == pytest.approx([0, 0, 0], abs=SLOP)


def test_point_to_body_distance_golden(world):
    sc, w, mug, palm = world
    _moved_mug(w, mug)
    d, cp, cn = w.point_distance(mug, [10, 11, 12])
    assert d == pytest.approx(33.8332, abs=SLOP)
    assert cn == pytest.approx([0.951057, -0.168109, 0.259287], abs=SLOP)
    assert cp == pytest.approx([42.1773, 5.31235, 20.7725], abs=SLOP)


def test_bounding_volumes_golden(world):
    sc, w, mug, palm = world
    bm = w.bounding_volumes(mug, 0)
    assert bm.shape == (1, 10)
    assert bm[0, 0:3] == pytest.approx([-11.662, -8.98983, 0.000190189], abs=SLOP)
    assert bm[0, 3:7] == pytest.approx([0.999999, -9.15765e-06, -1.47701e-06, 0.00157698], abs=SLOP)
    bp = w.bounding_volumes(palm, 0)
    assert bp.shape == (1, 10)
    assert bp[0, 0:3] == pytest.approx([0.013166, -9.91786, -38.5818], abs=SLOP)
    assert bp[0, 3:7] == pytest.approx([0.6664499, 0.236398, -0.66585779, -0.23789390], abs=SLOP)


def test_body_count_matches_primitive_planner_test():
    # test/primative_planner_tests.cpp:84: plannerMug has 10 bodies (mug + palm + 8 links)
    assert len(load_pack("barrett_mug").bodies) == 10


def test_transform_goldens():
    # test/transform_tests.cpp:16-43: point / normal transform values
    q = ref.transf_axis_angle(10.0, [1 / math.sqrt(3)] * 3)[0]
    m, q2 = ref.quat_roundtrip(q)
    # test/eigen_tests.cpp:9-32: quaternion -> matrix -> quaternion is the same rotation (sign may flip)
    assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12
    assert np.allclose(m @ m.T, np.eye(3), atol=1e-12)
    # composition order A_01 % (A_12 % A_23) == (A_01 % A_12) % A_23   (transform_tests.cpp:45-88)
    a = ref.transf_axis_angle(0.3, [1, 0, 0]); a = (a[0], np.array([1.0, 2, 3]))
    b = ref.transf_axis_angle(-1.1, [0, 1, 0]); b = (b[0], np.array([-4.0, 0.5, 2]))
    c = ref.transf_axis_angle(2.0, [0, 0, 1]); c = (c[0], np.array([0.0, 7, -1]))
    l = ref.transf_compose(a, ref.transf_compose(b, c))
    r = ref.transf_compose(ref.transf_compose(a, b), c)
    assert np.allclose(l[0], r[0], atol=1e-12) and np.allclose(l[1], r[1], atol=1e-10)
    p = np.array([0.3, -2.0, 5.0])
    assert np.allclose(ref.transf_apply(l, p), ref.transf_apply(a, ref.transf_apply(b, ref.transf_apply(c, p))), atol=1e-10)
    inv = ref.transf_inverse(l)
    assert np.allclose(ref.transf_apply(inv, ref.transf_apply(l, p)), p, atol=1e-10)


def test_transform_tests_values():
    # test/transform_tests.cpp: rotation of 90 deg about x then translation; values quoted in SURVEY 8(c)
    rot = ref.transf_axis_angle(-math.pi / 2, [1, 0, 0])
    assert rot[0] == pytest.approx([0.707107, -0.707107, 0, 0], abs=1e-6)
    j = ref.transf_jacobian((rot[0], np.array([1.0, 2.0, 3.0])))
    assert j.shape == (36,)
    assert j[3] == 0 and j[4] == 0 and j[5] == 0


def test_axis_angle_small_angle_snap():
    # src/transform.cpp:26-29: angle^2 < 1e-6 -> IDENTITY
    q, t = ref.transf_axis_angle(9.9e-4, [0, 0, 1])
    assert q == pytest.approx([1, 0, 0, 0], abs=0)
    q, t = ref.transf_axis_angle(1.1e-3, [0, 0, 1])
    assert q[3] > 0
