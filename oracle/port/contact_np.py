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


# ---- soft-finger contacts (SFCL) --------------------------------------------------------------------------------
# contactSetting.cpp:83-191 (neighbourhood radius, overlap, merge), softContact.cpp:69-117,204-358 (relPhi, relative
# curvature, mattress model, friction ellipsoid).  The surface fit itself is NOT restated: oracle.ref.soft_fit runs the
# reference's own FitParabola.h.  softContact.cpp is Qt/Coin-bound (restated); no reference test holds golden values for
# it, so relPhi / ellipse / wrench parity is pinned through the reference-run fit, the reference-run frame and the
# closed forms checked in tests/test_soft_contacts.py.

def soft_neighborhood_radius(y1, y2):
    rad = pow(1 / max(y1, y2), 0.333) * 1000.0 * 0.4
    if rad <= 3.0 and rad >= 10.0:
        rad = 5.0
    return rad


def _overlap(n1, n2):
    s = {tuple(p) for p in np.asarray(n2).reshape(-1, 3)}
    return any(tuple(p) in s for p in np.asarray(n1).reshape(-1, 3))


def _merge(n1, n2):
    out = [tuple(p) for p in np.asarray(n1).reshape(-1, 3)]
    for p in np.asarray(n2).reshape(-1, 3):
        if tuple(p) not in out:
            out.append(tuple(p))
    return np.array(out).reshape(-1, 3)


def contact_distance(row, tran1, tran2):
    from oracle.port import transf_np as T
    t1 = T.make(np.asarray(tran1[0])[None], np.asarray(tran1[1])[None]); t2 = T.make(np.asarray(tran2[0])[None], np.asarray(tran2[1])[None])
    p1 = T.apply(t1, np.asarray(row[0:3])[None])[0]; p2 = T.apply(t2, np.asarray(row[3:6])[None])[0]
    return float(np.linalg.norm(p1 - p2))


def merge_soft_neighborhoods(rows, tran1, tran2, nghbd1, nghbd2):
    rows = [np.array(r, dtype=np.float64) for r in rows]
    n1 = [np.asarray(a, dtype=np.float64).reshape(-1, 3) for a in nghbd1]
    n2 = [np.asarray(a, dtype=np.float64).reshape(-1, 3) for a in nghbd2]
    merged = True
    count = 0
    while merged:
        merged = False
        for r in range(len(rows)):
            for o in range(len(rows)):
                if o == r:
                    continue
                if _overlap(n1[r], n1[o]):
                    if contact_distance(rows[r], tran1, tran2) < contact_distance(rows[o], tran1, tran2):
                        keep, drop = r, o
                    else:
                        keep, drop = o, r
                    n1[keep] = _merge(n1[keep], n1[drop]); n2[keep] = _merge(n2[keep], n2[drop])
                    del rows[drop], n1[drop], n2[drop]
                    merged = True
                    count += 1
                    break
            if merged:
                break
    return np.array(rows).reshape(-1, 13), n1, n2, count


def soft_contact_side(pos, normal, nghbd, body_q):
    """per-side state of a SoftContact: frame, fit, main curvature direction in world coordinates"""
    from oracle.port import transf_np as T
    q, t, aff, n = contact_frame(pos, normal)
    fit, fit_rot = ref.soft_fit(pos, normal, nghbd)
    tb = T.make(np.asarray(body_q, dtype=np.float64)[None], np.zeros((1, 3)))
    r1dir = T.rotate(tb, (aff @ (fit_rot @ np.array([1.0, 0, 0])))[None])[0]
    return dict(aff=aff, n=n, fit=fit, dir=r1dir, pos=np.asarray(pos, dtype=np.float64))


def soft_contact_wrenches(this, mate, youngs, cof, cog, max_radius):
    """setUpFrictionEdges + computeWrenches of one SoftContact whose mate is `mate` (both from soft_contact_side)"""
    with np.errstate(all="ignore"):
        rel_phi = np.arccos((this["dir"] @ mate["dir"]) / (np.linalg.norm(this["dir"]) * np.linalg.norm(mate["dir"])))
        r1, r2 = this["fit"][3], this["fit"][4]
        m1, m2 = mate["fit"][3], mate["fit"][4]
        w = 0.0 if r1 < 0.0 else 1 / r1
        x = 0.0 if r2 < 0.0 else 1 / r2
        y = 0.0 if m1 < 0.0 else 1 / m1
        z = 0.0 if m2 < 0.0 else 1 / m2
        apb = 0.5 * (w + x + y + z)
        amb = 0.5 * np.sqrt((w - x) * (w - x) + (y - z) * (y - z) + 2 * np.cos(2 * rel_phi) * (w - x) * (y - z))
        r1p = r2p = 0.0
        if not (amb > apb):
            r1p = 1 / (apb + amb) if apb + amb != 0 else -1.0
            r2p = -1.0 if apb == amb else 1 / (apb - amb)
        if r1p < 0:
            r1p = 20
        if r2p < 0:
            r2p = 20
        h = 0.003
        delta = np.sqrt(5 * h / (youngs * np.pi * np.sqrt(r1p * 0.001 * r2p * 0.001)))
        major = np.sqrt(2 * delta * r1p * 0.001); minor = np.sqrt(2 * delta * r2p * 0.001)
        tdn = 8 * np.sqrt(major * minor) / 15
        edges = friction_ellipsoid([1, 5, 8, 5, 1], [np.pi / 2, np.pi / 2 * 0.5, 0.0, -np.pi / 2 * 0.5, -np.pi / 2], [1, 1, tdn * 1000])
    tx, ty, n = this["aff"][:, 0], this["aff"][:, 1], this["n"]
    radius = this["pos"] - np.asarray(cog, dtype=np.float64)
    out = []
    for e in edges:
        force = n + tx * cof * e[0] + ty * cof * e[1]
        torque = (np.cross(radius, force) + n * cof * e[5]) / max_radius
        out.append(np.concatenate([force, torque]))
    return np.array(out), np.array([major, minor, rel_phi, r1p, r2p, tdn])
