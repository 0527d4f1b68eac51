"""CPU: contact set-up pinned to the reference's OWN lines.

oracle/_ref/libgraspit_ref_contact.so holds Contact / PointContact / SoftContact member functions and the free functions
of contactSetting.cpp cut out of the reference's sources at build time (oracle/ref_contact.cpp, oracle/Makefile) and
compiled against the reference's own contact headers.  Both the product's host utilities (libb2c.so, through
graspit_b200.engine) and the numpy restatement the other tests use (oracle/port/contact_np.py) are held to it here."""
import os

import numpy as np
import pytest

from graspit_b200 import engine as E
from oracle import ref
from oracle.port import contact_np as cn

pytestmark = pytest.mark.skipif(not os.path.exists(ref.CONTACT_LIB_PATH), reason="oracle/_ref/libgraspit_ref_contact.so not built")


def _rows(n, seed):
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 13))
    rows[:, 0:3] = rng.uniform(-80, 80, (n, 3)); rows[:, 3:6] = rng.uniform(-80, 80, (n, 3))
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    rows[:, 6:9] = nrm; rows[:, 9:12] = -nrm
    # the special cases of Contact::Contact: normal along +/- x, an un-normalised normal, a normal just off the x axis
    rows[0, 6:9] = [1, 0, 0]; rows[1, 6:9] = [-1, 0, 0]; rows[2, 6:9] = [0, 0, 2.5]; rows[3, 6:9] = [1 - 1e-12, 1e-7, 0]
    return rows


def _tran(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return q, rng.uniform(-50, 50, 3)


def _same_rotation(qa, qb, tol=1e-12):
    return min(np.abs(qa - qb).max(), np.abs(qa + qb).max()) < tol


def test_point_contact_frames_edges_and_wrenches():
    rows = _rows(120, 11)
    cog, max_radius = np.array([3.0, -2.0, 10.0]), 85.0
    for cof in (0.0, 0.5, 1.2):
        frames = E.contact_frames(rows, 1)
        wr = E.point_contact_wrenches(rows, 1, cof, cog, max_radius)
        for i, r in enumerate(rows):
            f_ref, e_ref, w_ref = ref.point_contact(r[0:3], r[6:9], cof, cog, max_radius)
            assert len(w_ref) == 8
            assert np.allclose(frames[i, 4:], f_ref[4:], rtol=0, atol=1e-12) and _same_rotation(frames[i, :4], f_ref[:4])
            assert np.allclose(wr[i], w_ref, rtol=1e-12, atol=1e-12), (i, cof)
            # the restatement used by the other tests
            assert np.allclose(cn.point_contact_wrenches(r[0:3], r[6:9], cof, cog, max_radius), w_ref, rtol=1e-12, atol=1e-12)
            assert np.allclose(cn.friction_ellipsoid([8], [0.0], [1, 1, 1]), e_ref, rtol=0, atol=1e-15)


def test_check_contact_normals():
    rng = np.random.default_rng(12)
    rows = _rows(80, 13)
    rows[::3, 6:9] *= -1
    rows[1::4, 9:12] *= -1
    t1, t2 = _tran(rng), _tran(rng)
    out, nflip = E.check_contact_normals(rows, t1, t2)
    bad = 0
    for i, r in enumerate(rows):
        rr, ok = ref.check_contact_normals_ref(r, t1, t2)
        bad += not ok
        assert np.array_equal(out[i, :12], rr), i
        pr, pok = cn.check_contact_normals(r, t1, t2)
        assert np.array_equal(np.asarray(pr)[:12], rr) and pok == ok
    assert nflip == bad and 0 < bad < len(rows)


def test_soft_neighborhood_radius():
    for y1, y2 in ((1.5e6, -1.0), (-1.0, 1.5e6), (2.0e5, 8.0e6), (1.0e3, 1.0e3), (5.0e7, 1.0)):
        want = ref.soft_neighborhood_radius_ref(y1, y2)
        assert E.soft_neighborhood_radius(y1, y2) == pytest.approx(want, rel=1e-14)
        assert cn.soft_neighborhood_radius(y1, y2) == pytest.approx(want, rel=1e-14)


def test_merge_soft_neighborhoods():
    rng = np.random.default_rng(14)
    grid = np.round(rng.uniform(-20, 20, (60, 3)), 3)            # shared mesh vertices
    t1, t2 = _tran(rng), _tran(rng)
    merged_any = 0
    for trial in range(40):
        n = int(rng.integers(1, 10))
        rows = np.zeros((n, 13))
        rows[:, 0:3] = rng.uniform(-20, 20, (n, 3)); rows[:, 3:6] = rng.uniform(-20, 20, (n, 3))
        rows[:, 6:9] = rng.normal(size=(n, 3)); rows[:, 9:12] = -rows[:, 6:9]
        n1 = [grid[rng.choice(60, int(rng.integers(1, 8)), replace=False)] for _ in range(n)]
        n2 = [grid[rng.choice(60, int(rng.integers(0, 8)), replace=False)] for _ in range(n)]
        keep, a_ref, b_ref = ref.merge_soft_neighborhoods_ref(rows, t1, t2, n1, n2)
        r_e, a_e, b_e, m_e = E.soft_contact_merge(rows, t1, t2, n1, n2)
        assert m_e == n - len(keep) and len(r_e) == len(keep)
        merged_any += m_e > 0
        assert np.array_equal(r_e[:, :12], rows[keep, :12])                       # same survivors, same order
        for x, y in zip(a_e + b_e, a_ref + b_ref):
            assert np.array_equal(np.asarray(x).reshape(-1, 3), np.asarray(y).reshape(-1, 3))
        r_o, a_o, b_o, m_o = cn.merge_soft_neighborhoods(rows, t1, t2, n1, n2)
        assert m_o == m_e and np.array_equal(np.asarray(r_o)[:, :12], rows[keep, :12])
    assert merged_any >= 10


def _patch(rng, pos, normal, r1, r2, ang, k=40, size=3.0):
    """k points of the paraboloid z = x^2/(2 r1) + y^2/(2 r2) rotated by ang about the normal, laid out so that the
    reference's `frame.inverse().affine() * point` (rotation only, softContact.cpp:42-44) sees them in the contact frame"""
    _, _, aff, _ = cn.contact_frame(pos, normal)
    xy = rng.uniform(-size, size, (k, 2))
    c, s = np.cos(ang), np.sin(ang)
    u = c * xy[:, 0] + s * xy[:, 1]; v = -s * xy[:, 0] + c * xy[:, 1]
    z = (u * u / (2 * r1) if r1 > 0 else 0 * u) + (v * v / (2 * r2) if r2 > 0 else 0 * v)
    return np.stack([xy[:, 0], xy[:, 1], z], axis=1) @ aff.T


def test_soft_contact_pair_fit_ellipse_and_wrenches():
    rng = np.random.default_rng(15)
    n = 80
    rows, n1, n2 = np.zeros((n, 13)), [], []
    for i in range(n):
        pos1 = rng.uniform(-30, 30, 3); pos2 = rng.uniform(-30, 30, 3); nrm = rng.normal(size=3)
        rows[i, 0:3] = pos1; rows[i, 3:6] = pos2; rows[i, 6:9] = nrm; rows[i, 9:12] = -nrm
        n1.append(_patch(rng, pos1, nrm, rng.uniform(4, 40), rng.uniform(4, 40), rng.uniform(-1, 1)))
        flat = i % 4 == 0
        n2.append(_patch(rng, pos2, -nrm, -1 if flat else rng.uniform(4, 40), -1 if flat else rng.uniform(4, 40), rng.uniform(-1, 1)))
    t1, t2 = _tran(rng), _tran(rng)
    cog, max_r, cof, y1, y2 = np.array([1.0, 2.0, 3.0]), 70.0, 0.8, 1.5e6, 3.0e5
    f1 = E.soft_contact_fit(rows, 1, n1); f2 = E.soft_contact_fit(rows, 2, n2)
    w1, e1 = E.soft_contact_wrenches(rows, 1, f1, f2, t1[0], t2[0], max(y1, y2), cof, cog, max_r)
    nvalid = 0
    for i, r in enumerate(rows):
        w_ref, i1, i2 = ref.soft_contact_pair(t1, t2, r[0:3], r[6:9], n1[i], r[3:6], r[9:12], n2[i], y1, y2, cof, cog, max_r)
        assert len(w_ref) == 20
        # the surface fits (a, b, c, r1, r2) of both sides
        assert np.allclose(f1[i, :5], i1[5:10], rtol=1e-9, atol=1e-12, equal_nan=True), (i, f1[i], i1)
        assert np.allclose(f2[i, :5], i2[5:10], rtol=1e-9, atol=1e-12, equal_nan=True)
        # contact ellipse: major, minor, relPhi, r1', r2' (torque / N is what setUpFrictionEdges derives from them)
        assert np.allclose(e1[i, :5], i1[:5], rtol=1e-6, atol=1e-12, equal_nan=True), (i, e1[i], i1[:5])
        assert np.allclose(w1[i], w_ref, rtol=1e-6, atol=1e-9, equal_nan=True), i
        nvalid += bool(np.isfinite(i1[:5]).all())
        # the mate carries the pair's ellipse too (softContact.cpp:345-347)
        assert np.allclose(i1[[0, 1, 2, 3, 4]], i2[[0, 1, 2, 3, 4]], equal_nan=True)
        # and the restatement
        a = cn.soft_contact_side(r[0:3], r[6:9], n1[i], t1[0]); b = cn.soft_contact_side(r[3:6], r[9:12], n2[i], t2[0])
        wa, ea = cn.soft_contact_wrenches(a, b, max(y1, y2), cof, cog, max_r)
        assert np.allclose(wa, w_ref, rtol=1e-6, atol=1e-9, equal_nan=True)
    assert nvalid > n // 2
