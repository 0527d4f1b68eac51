"""CPU: soft-finger (SFCL) contact utilities of libb2c.so -- neighbourhood merge, paraboloid fit, relative curvature,
mattress-model ellipse, 20-edge friction ellipsoid wrenches -- against the reference's own FitParabola.h (run through
oracle/_ref) and the numpy restatement of softContact.cpp / contactSetting.cpp (oracle/port/contact_np.py)."""
import numpy as np

from graspit_b200 import engine as E
from oracle import ref
from oracle.port import contact_np as cn


def _row(p1, n1, p2=None):
    r = np.zeros(13)
    n1 = np.asarray(n1, dtype=np.float64)
    r[0:3] = p1; r[3:6] = p1 if p2 is None else p2
    r[6:9] = n1; r[9:12] = -n1
    return r


def _patch(rng, pos, normal, r1, r2, ang, k=40, size=3.0):
    """k points of the paraboloid z = x^2/(2 r1) + y^2/(2 r2), rotated by ang about the normal, in the contact frame
    of (pos, normal), expressed in the body frame"""
    _, _, aff, _ = cn.contact_frame(pos, normal)
    xy = rng.uniform(-size, size, (k, 2))
    c, s = np.cos(ang), np.sin(ang)
    u = c * xy[:, 0] + s * xy[:, 1]; v = -s * xy[:, 0] + c * xy[:, 1]
    z = (u * u / (2 * r1) if r1 > 0 else 0 * u) + (v * v / (2 * r2) if r2 > 0 else 0 * v)
    local = np.stack([xy[:, 0], xy[:, 1], z], axis=1)
    return local @ aff.T          # the reference rotates back with frame^-1 only (no translation), so neither do we


def _cond(pos, normal, pts):
    if len(pts) < 3:
        return 1.0
    _, _, aff, _ = cn.contact_frame(pos, normal)
    p = np.asarray(pts) @ aff
    x, y = p[:, 0], p[:, 1]
    B = np.stack([x * x, y * y, x * y], axis=1)
    return float(np.linalg.cond(B.T @ B))


def test_neighborhood_radius():
    # rubber, E = 1.5 MPa: (1/1.5e6)^0.333 * 400 = 3.52 mm (contactSetting.cpp:129)
    assert abs(E.soft_neighborhood_radius(1.5e6, 1.0e5) - (1 / 1.5e6) ** 0.333 * 400.0) < 1e-12
    assert abs(E.soft_neighborhood_radius(1.5e6, 0) - cn.soft_neighborhood_radius(1.5e6, 0)) < 1e-12
    assert 3.4 < E.soft_neighborhood_radius(0, 1.5e6) < 3.6


def test_fit_recovers_radii_of_curvature():
    rng = np.random.default_rng(5)
    # axis-aligned paraboloid at the origin of the body frame, normal +z: r1 = 12, r2 = 30
    pts = _patch(rng, [0, 0, 0], [0, 0, 1.0], 12.0, 30.0, 0.0)
    f = E.soft_contact_fit(_row([0, 0, 0], [0, 0, 1.0])[None], 1, [pts])[0]
    assert abs(f[0] - 1 / 24.0) < 1e-9 and abs(f[1] - 1 / 60.0) < 1e-9 and f[2] == 0.0
    assert abs(f[3] - 12.0) < 1e-6 and abs(f[4] - 30.0) < 1e-6 and f[5] == 0.0
    # plane: every coefficient snaps to 0, radii -1 (flat)
    flat = _patch(rng, [0, 0, 0], [0, 0, 1.0], -1, -1, 0.0)
    f = E.soft_contact_fit(_row([0, 0, 0], [0, 0, 1.0])[None], 1, [flat])[0]
    assert (f[:3] == 0).all() and f[3] == -1.0 and f[4] == -1.0
    # rotated principal axes: the mixed term is removed by rot_angle, radii are recovered (in either order)
    pts = _patch(rng, [0, 0, 0], [0.3, -0.2, 0.9], 8.0, 25.0, 0.6)
    f = E.soft_contact_fit(_row([0, 0, 0], [0.3, -0.2, 0.9])[None], 1, [pts])[0]
    assert abs(f[2]) > 1e-3 and f[5] != 0.0
    assert np.allclose(sorted([f[3], f[4]]), [8.0, 25.0], rtol=1e-6)


def test_fit_matches_reference_fitparabola():
    rng = np.random.default_rng(6)
    rows, ngh = [], []
    for i in range(120):
        pos = rng.uniform(-40, 40, 3) if i % 3 else rng.uniform(-2, 2, 3)
        nrm = rng.normal(size=3)
        if i == 0:
            nrm = np.array([1.0, 0, 0])
        if i == 1:
            nrm = np.array([-2.0, 0, 0])
        rows.append(_row(pos, nrm))
        # real neighbourhoods sit around the contact position (and are not re-centred by the reference)
        ngh.append(pos + _patch(rng, pos, nrm, rng.uniform(4, 40), rng.uniform(4, 40), rng.uniform(-1, 1), k=int(rng.integers(3, 60))))
    ngh[5] = ngh[5][:0]           # empty neighbourhood: singular normal equations, as in the reference
    got = E.soft_contact_fit(np.array(rows), 1, ngh)
    tight = 0
    for i, r in enumerate(rows):
        want, _ = ref.soft_fit(r[0:3], r[6:9], ngh[i])
        # the reference does not re-centre the neighbourhood, so its 3x3 normal equations are ill-conditioned for
        # contacts far from the body origin: one ulp in the rotated points moves the solution by cond(A) ulps.  The
        # tolerance follows the conditioning of each case; well-conditioned cases must agree to 1e-9.
        rtol = max(1e-9, 1e-14 * _cond(r[0:3], r[6:9], ngh[i]))
        tight += rtol == 1e-9
        assert rtol < 1e-2 or len(ngh[i]) < 3
        assert np.allclose(got[i], want, rtol=rtol, atol=1e-12, equal_nan=True), (i, got[i], want, rtol)
    assert tight >= 30
    # side 2 reads b2_pos / b2_normal
    got2 = E.soft_contact_fit(np.array(rows[:10]), 2, ngh[:10])
    for i, r in enumerate(rows[:10]):
        want, _ = ref.soft_fit(r[3:6], r[9:12], ngh[i])
        assert np.allclose(got2[i], want, rtol=max(1e-9, 1e-14 * _cond(r[3:6], r[9:12], ngh[i])), atol=1e-12, equal_nan=True)


def test_sphere_on_plane_contact_ellipse():
    # ball of radius 10 mm (r1 = r2 = 10) on a flat mate (-1, -1): relative radii 10, 10 -> circular contact patch
    row = _row([0, 0, 0], [0, 0, 1.0])[None]
    fit_ball = np.array([[0.05, 0.05, 0.0, 10.0, 10.0, 0.0]]); fit_flat = np.array([[0, 0, 0, -1.0, -1.0, 0.0]])
    q = np.array([1.0, 0, 0, 0])
    w, ell = E.soft_contact_wrenches(row, 1, fit_ball, fit_flat, q, q, 1.5e6, 1.0, [0, 0, -10.0], 50.0)
    delta = np.sqrt(5 * 0.003 / (1.5e6 * np.pi * 0.010))
    a = np.sqrt(2 * delta * 0.010)
    # the mate's frame is built from the opposite normal, so its tangentX is reversed: relPhi = pi (cos 2 relPhi = 1)
    assert np.allclose(ell[0], [a, a, np.pi, 10.0, 10.0, 8 * a / 15], rtol=1e-12)
    assert w.shape == (1, 20, 6)
    # pole edges carry pure torque about the normal: force = normal, torque = +- cof * tdn * 1000 / max_radius
    assert np.allclose(w[0, 0, :3], [0, 0, 1]) and np.allclose(w[0, 19, :3], [0, 0, 1])
    assert np.allclose(w[0, 0, 3:], [0, 0, 8 * a / 15 * 1000 / 50.0]) and np.allclose(w[0, 19, 3:], [0, 0, -8 * a / 15 * 1000 / 50.0])
    # equator edges: the PCWF circle
    tang = w[0, 6:14, :2]
    assert np.allclose(np.linalg.norm(tang, axis=1), 1.0)
    # two flat bodies: degenerate curvature -> both relative radii fall back to 20 mm
    _, ell = E.soft_contact_wrenches(row, 1, fit_flat, fit_flat, q, q, 1.5e6, 1.0, [0, 0, 0], 50.0)
    assert ell[0, 3] == 20 and ell[0, 4] == 20


def test_wrenches_match_restatement():
    rng = np.random.default_rng(8)
    n = 60
    rows, n1, n2 = [], [], []
    for i in range(n):
        pos1 = rng.uniform(-30, 30, 3); pos2 = rng.uniform(-30, 30, 3); nrm = rng.normal(size=3)
        rows.append(_row(pos1, nrm, pos2))
        n1.append(pos1 + _patch(rng, pos1, nrm, rng.uniform(4, 40), rng.uniform(4, 40), rng.uniform(-1, 1)))
        flat = i % 4 == 0
        n2.append(pos2 + _patch(rng, pos2, -nrm, -1 if flat else rng.uniform(4, 40), -1 if flat else rng.uniform(4, 40), rng.uniform(-1, 1)))
    rows = np.array(rows)
    qa = rng.normal(size=4); qa /= np.linalg.norm(qa); qb = rng.normal(size=4); qb /= np.linalg.norm(qb)
    f1 = E.soft_contact_fit(rows, 1, n1); f2 = E.soft_contact_fit(rows, 2, n2)
    cog, max_r, cof, youngs = np.array([1.0, 2.0, 3.0]), 70.0, 0.8, 1.5e6
    w1, e1 = E.soft_contact_wrenches(rows, 1, f1, f2, qa, qb, youngs, cof, cog, max_r)
    w2, e2 = E.soft_contact_wrenches(rows, 2, f2, f1, qb, qa, youngs, cof, cog, max_r)
    nvalid = 0
    for i, r in enumerate(rows):
        a = cn.soft_contact_side(r[0:3], r[6:9], n1[i], qa)
        b = cn.soft_contact_side(r[3:6], r[9:12], n2[i], qb)
        wa, ea = cn.soft_contact_wrenches(a, b, youngs, cof, cog, max_r)
        wb, eb = cn.soft_contact_wrenches(b, a, youngs, cof, cog, max_r)
        assert np.allclose(e1[i], ea, rtol=1e-6, atol=1e-12, equal_nan=True), (i, e1[i], ea)
        assert np.allclose(e2[i], eb, rtol=1e-6, atol=1e-12, equal_nan=True)
        assert np.allclose(w1[i], wa, rtol=1e-6, atol=1e-9, equal_nan=True)
        assert np.allclose(w2[i], wb, rtol=1e-6, atol=1e-9, equal_nan=True)
        if np.isfinite(ea).all():
            nvalid += 1
            # the ellipse is a property of the pair: both sides agree (softContact.cpp:345-347)
            assert np.allclose(ea[[0, 1, 3, 4, 5]], eb[[0, 1, 3, 4, 5]], rtol=1e-9)
    assert nvalid > n // 2


def test_merge_matches_restatement():
    rng = np.random.default_rng(9)
    grid = np.round(rng.uniform(-20, 20, (60, 3)), 3)            # shared mesh vertices
    q1 = rng.normal(size=4); q1 /= np.linalg.norm(q1); q2 = rng.normal(size=4); q2 /= np.linalg.norm(q2)
    t1, t2 = rng.uniform(-10, 10, 3), rng.uniform(-10, 10, 3)
    for trial in range(20):
        n = int(rng.integers(1, 9))
        rows = np.array([_row(rng.uniform(-20, 20, 3), rng.normal(size=3), rng.uniform(-20, 20, 3)) for _ in range(n)])
        n1 = [grid[rng.choice(60, int(rng.integers(1, 8)), replace=False)] for _ in range(n)]
        n2 = [grid[rng.choice(60, int(rng.integers(0, 8)), replace=False)] for _ in range(n)]
        r_e, a_e, b_e, m_e = E.soft_contact_merge(rows, (q1, t1), (q2, t2), n1, n2)
        r_o, a_o, b_o, m_o = cn.merge_soft_neighborhoods(rows, (q1, t1), (q2, t2), n1, n2)
        assert m_e == m_o and len(r_e) == len(r_o) == n - m_o
        assert np.array_equal(r_e, r_o)
        for x, y in zip(a_e + b_e, a_o + b_o):
            assert np.array_equal(np.asarray(x).reshape(-1, 3), np.asarray(y).reshape(-1, 3))
        # survivors no longer overlap on body 1
        for i in range(len(a_e)):
            for j in range(i + 1, len(a_e)):
                assert not ({tuple(p) for p in a_e[i]} & {tuple(p) for p in a_e[j]})
    # nothing to merge: untouched
    rows = np.array([_row([0, 0, 0], [0, 0, 1.0]), _row([5, 0, 0], [0, 0, 1.0])])
    r, a, b, m = E.soft_contact_merge(rows, (q1, t1), (q2, t2), [grid[:3], grid[3:6]], [grid[:2], grid[:2]])
    assert m == 0 and np.array_equal(r, rows) and np.array_equal(a[1], grid[3:6])
