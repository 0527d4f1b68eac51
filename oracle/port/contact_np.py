"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp64) of the reference's contact frame and point-contact
friction cone: Contact::Contact (src/contact/contact.cpp:67-99), PointContact::setUpFrictionEdges
(src/contact/pointContact.cpp:44-56), Contact::setUpFrictionEllipsoid (contact.cpp:180-216),
Contact::wrenchFromFrictionEdge / computeWrenches (contact.cpp:124-165), checkContactNormals
(src/contactSetting.cpp:54-81).

contact.cpp is Qt/Coin-bound and cannot be compiled here, so these formulas are restated; the frame itself is built by
the reference's own transf::COORDINATE (compiled verbatim in oracle/_ref), which pins the quaternion convention.
The reference holds no golden vectors for friction cones: wrench parity is pinned only through that frame and the
closed forms checked by hand in tests/test_contact_cones.py.
"""
from __future__ import annotations

import numpy as np

from oracle import ref

MACHINE_ZERO = 1.0e-9      # include/graspit/matvec3D.h:53


def _normalized(v):
    n2 = float(v @ v)
    return v / np.sqrt(n2) if n2 > 0 else v


def contact_frame(pos, normal):
    normal = _normalized(np.asarray(normal, dtype=np.float64))
    if abs(normal @ np.array([1.0, 0, 0])) > 1.0 - MACHINE_ZERO:
        tx = _normalized(np.cross(normal, [0, 0, 1.0]))
    else:
        tx = _normalized(np.cross(normal, [1.0, 0, 0]))
    ty = _normalized(np.cross(normal, tx))
    q, t, aff = ref.transf_coordinate(pos, tx, ty)
    return q, t, aff, normal


def friction_ellipsoid(num_dirs, phi, eccen):
    edges = []
    for nd, ph in zip(num_dirs, phi):
        cosphi, sinphi = np.cos(ph), np.sin(ph)
        for j in range(nd):
            theta = j * 2 * np.pi / nd
            num = np.cos(theta) * cosphi
            denom = num * num / (eccen[0] * eccen[0])
            num = np.sin(theta) * cosphi
            denom += num * num / (eccen[1] * eccen[1])
            num = sinphi
            denom += num * num / (eccen[2] * eccen[2])
            denom = np.sqrt(denom)
            edges.append([np.cos(theta) * cosphi / denom, np.sin(theta) * cosphi / denom, 0, 0, 0, sinphi / denom])
    return np.array(edges)


def point_contact_wrenches(pos, normal, cof, cog, max_radius):
    q, t, aff, n = contact_frame(pos, normal)
    tx, ty = aff[:, 0], aff[:, 1]
    edges = friction_ellipsoid([8], [0.0], [1.0, 1.0, 1.0])
    radius = np.asarray(pos, dtype=np.float64) - np.asarray(cog, dtype=np.float64)
    out = []
    for e in edges:
        force = n + tx * cof * e[0] + ty * cof * e[1]
        torque = (np.cross(radius, force) + n * cof * e[5]) / max_radius
        out.append(np.concatenate([force, torque]))
    return np.array(out)


def check_contact_normals(row, tran1, tran2):
    """row: 13 doubles (b1_pos, b2_pos, b1_normal, b2_normal, distSq); returns (row', ok)"""
    from oracle.port import transf_np as T
    r = np.array(row, dtype=np.float64)
    t1 = T.make(np.asarray(tran1[0])[None], np.asarray(tran1[1])[None]); t2 = T.make(np.asarray(tran2[0])[None], np.asarray(tran2[1])[None])
    n1 = T.rotate(t1, r[6:9][None])[0]; n2 = T.rotate(t2, r[9:12][None])[0]
    p1 = T.apply(t1, r[0:3][None])[0]; p2 = T.apply(t2, r[3:6][None])[0]
    d = p2 - p1
    ok = True
    if d @ n1 > 0:
        ok = False; r[6:9] = -r[6:9]
    if d @ n2 < 0:
        ok = False; r[9:12] = -r[9:12]
    return r, ok
