"""CPU: contact frames, point-contact friction cones and checkContactNormals (host utilities of libb2c.so) against the
numpy restatement whose frames come from the reference's own transf::COORDINATE (oracle/_ref)."""
import numpy as np

from graspit_b200 import engine
from oracle.port import contact_np as cn


def _rows(n, seed):
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 13))
    rows[:, 0:3] = rng.uniform(-80, 80, (n, 3)); rows[:, 3:6] = rng.uniform(-80, 80, (n, 3))
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    rows[:, 6:9] = nrm; rows[:, 9:12] = -nrm
    # the special cases of Contact::Contact: normal along +/- x, and an un-normalised normal
    rows[0, 6:9] = [1, 0, 0]; rows[1, 6:9] = [-1, 0, 0]; rows[2, 6:9] = [0, 0, 2.5]; rows[3, 6:9] = [1 - 1e-12, 1e-7, 0]
    return rows


def test_friction_circle_closed_form():
    e = cn.friction_ellipsoid([8], [0.0], [1, 1, 1])
    assert e.shape == (8, 6)
    ang = np.arange(8) * np.pi / 4
    assert np.allclose(e[:, 0], np.cos(ang)) and np.allclose(e[:, 1], np.sin(ang)) and np.allclose(e[:, 2:], 0)
    # soft-finger style sample: latitude 45 deg on an ellipsoid with eccentricities (1, 1, 2): x = cos45/d, z = sin45/d
    e2 = cn.friction_ellipsoid([1], [np.pi / 4], [1, 1, 2])
    d = np.sqrt(0.5 + 0.5 / 4)
    assert np.allclose(e2[0], [np.sqrt(0.5) / d, 0, 0, 0, 0, np.sqrt(0.5) / d])


def test_contact_frames_match_reference_coordinate_frames():
    rows = _rows(200, 1)
    for side in (1, 2):
        f = engine.contact_frames(rows, side)
        for i, r in enumerate(rows):
            pos = r[0:3] if side == 1 else r[3:6]
            nrm = r[6:9] if side == 1 else r[9:12]
            q, t, aff, n = cn.contact_frame(pos, nrm)
            assert np.allclose(f[i, 4:], t)
            assert min(np.abs(f[i, :4] - q).max(), np.abs(f[i, :4] + q).max()) < 1e-12
            # z axis of the frame is the (normalised) contact normal
            assert np.allclose(aff[:, 2], n, atol=1e-12)


def test_point_contact_wrenches():
    rows = _rows(100, 2)
    cog, max_radius, cof = np.array([3.0, -2.0, 10.0]), 85.0, 0.5
    w = engine.point_contact_wrenches(rows, 1, cof, cog, max_radius)
    assert w.shape == (100, 8, 6)
    for i, r in enumerate(rows):
        ref_w = cn.point_contact_wrenches(r[0:3], r[6:9], cof, cog, max_radius)
        assert np.allclose(w[i], ref_w, rtol=1e-12, atol=1e-12)
        n = r[6:9] / np.linalg.norm(r[6:9])
        # every boundary force lies on the friction cone: normal component 1, tangential magnitude cof
        assert np.allclose(w[i, :, :3] @ n, 1.0)
        tang = w[i, :, :3] - np.outer(w[i, :, :3] @ n, n)
        assert np.allclose(np.linalg.norm(tang, axis=1), cof)
    # frictionless: all eight wrenches are the normal force
    w0 = engine.point_contact_wrenches(rows[:3], 1, 0.0, cog, max_radius)
    assert np.allclose(w0[:, :, :3], (rows[:3, 6:9] / np.linalg.norm(rows[:3, 6:9], axis=1, keepdims=True))[:, None, :])


def test_check_contact_normals():
    rng = np.random.default_rng(3)
    rows = _rows(60, 4)
    rows[::3, 6:9] *= -1          # some normals point the wrong way
    rows[1::4, 9:12] *= -1
    q1 = rng.normal(size=4); q1 /= np.linalg.norm(q1); q2 = rng.normal(size=4); q2 /= np.linalg.norm(q2)
    t1, t2 = rng.uniform(-50, 50, 3), rng.uniform(-50, 50, 3)
    out, nflip = engine.check_contact_normals(rows, (q1, t1), (q2, t2))
    want, bad = [], 0
    for r in rows:
        rr, ok = cn.check_contact_normals(r, (q1, t1), (q2, t2))
        want.append(rr); bad += (not ok)
    assert np.allclose(out, np.array(want)) and nflip == bad and 0 < bad < 60
    # idempotent
    out2, nflip2 = engine.check_contact_normals(out, (q1, t1), (q2, t2))
    assert nflip2 == 0 and np.allclose(out2, out)
