"""Asset-format readers (Inventor / VRML scene graphs, GraspIt! XML, scene packs).  CPU only; uses inline text
because GraspIt!'s asset tree is not available where the tests run."""
import os

import numpy as np

from graspit_b200.formats import transf as tf
from graspit_b200.formats.scenegraph import triangles_from_text
from graspit_b200.formats.scenepack import load_pack, save_pack
from graspit_b200.formats.xmlmodel import parse_transform

IV_FACESET = """#Inventor V2.0 ascii
Separator {
  Translation { translation 1 2 3 }
  Coordinate3 { point [ 0 0 0, 1 0 0, 1 1 0, 0 1 0,  0 0 1, 1 0 1, 0 1 1,  5 5 5, 5 5 5, 6 5 5 ] }
  FaceSet { numVertices [ 4, 3, 3 ] }
}
"""

IV_STRIP = """#Inventor V2.1 ascii
Separator {
    RotationXYZ { axis Z  angle 1.5707963 }
    IndexedTriangleStripSet {
      vertexProperty VertexProperty { vertex [ 0 0 0, 1 0 0, 0 1 0, 1 1 0, 0 2 0 ] }
      coordIndex [ 0, 1, 2, 3, 4, -1 ]
    }
}
"""

VRML2 = """#VRML V2.0 utf8
DEF Scene Group { children [
  Transform { scale 2 2 2 children [
    Transform { translation 1 0 0 children [
      Shape { appearance Appearance { material Material { } }
        geometry IndexedFaceSet { coord Coordinate { point [ 0 0 0  1 0 0  1 1 0  0 1 0 ] }
                                  coordIndex [ 0 1 2 3 -1 ] } } ] } ] } ] }
"""


def test_faceset_quads_and_degenerates():
    t = triangles_from_text(IV_FACESET)
    assert t.dtype == np.float32 and t.shape == (3, 3, 3)      # quad -> 2 triangles, 1 triangle, degenerate dropped
    assert np.allclose(t[0], [[1, 2, 3], [2, 2, 3], [2, 3, 3]])
    assert np.allclose(t[1], [[1, 2, 3], [2, 3, 3], [1, 3, 3]])


def test_triangle_strip_and_rotation():
    t = triangles_from_text(IV_STRIP)
    assert t.shape == (3, 3, 3)
    # rotated 90 deg about z: (1,0,0) -> (0,1,0)
    assert np.allclose(t[0][1], [0, 1, 0], atol=1e-6)
    # alternate strip triangles are flipped to keep the orientation
    n0 = np.cross(t[0][1] - t[0][0], t[0][2] - t[0][0]); n1 = np.cross(t[1][1] - t[1][0], t[1][2] - t[1][0])
    assert np.dot(n0, n1) > 0


def test_vrml2_nested_transforms():
    t = triangles_from_text(VRML2)
    assert t.shape == (2, 3, 3)
    assert np.allclose(t.reshape(-1, 3).min(0), [2, 0, 0]) and np.allclose(t.reshape(-1, 3).max(0), [4, 2, 0])


def test_transform_element_composes_left_to_right():
    import xml.etree.ElementTree as ET
    e = ET.fromstring("<transform><translation>25.0 0 -1.0</translation><rotation>180 y</rotation></transform>")
    q, t = parse_transform(e)
    assert np.allclose(t, [25, 0, -1]) and np.allclose(np.abs(q), [0, 0, 1, 0], atol=1e-12)
    e = ET.fromstring("<transform><fullTransform>(+1 0 0 0)[+8.37264e-06 +0.781957 +105.818]</fullTransform></transform>")
    q, t = parse_transform(e)
    assert np.allclose(q, [1, 0, 0, 0]) and np.allclose(t, [8.37264e-06, 0.781957, 105.818])
    e = ET.fromstring("<transform><rotation>-90 y</rotation><rotation>-90 z</rotation></transform>")
    q, _ = parse_transform(e)
    a = tf.compose(tf.angle_axis(-np.pi / 2, (0, 1, 0)), tf.angle_axis(-np.pi / 2, (0, 0, 1)))
    assert np.allclose(q, a[0])


def test_scene_pack_contents_and_roundtrip(tmp_path):
    sc = load_pack("barrett_mug")
    assert [len(b.tris) for b in sc.bodies] == [3450, 268, 60, 72, 52, 60, 72, 52, 72, 52]
    r = sc.robots[0]
    assert r.n_dof == 4 and len(r.vcontacts) == 19 and r.eg.shape == (2, 4)
    assert len(sc.disabled_pairs) == 10
    assert np.allclose(r.dof_max, [np.pi, 2.513274, 2.513274, 2.513274], atol=1e-5)
    p = os.path.join(tmp_path, "x.npz")
    save_pack(sc, p)
    sc2 = load_pack(p)
    assert all(np.array_equal(a.tris, b.tris) for a, b in zip(sc.bodies, sc2.bodies))
    assert sc2.disabled_pairs == sc.disabled_pairs
    assert np.allclose(sc2.robots[0].chains[2].tran[0], r.chains[2].tran[0])
    assert [j.ratio for j in sc2.robots[0].chains[0].joints] == [1.0, 1.0, 0.333333333]
