"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libgraspit_ref.so.

The library is the reference's own GraspitCollision backend compiled verbatim (see
oracle/Makefile, oracle/ref_driver.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module; the product path
(graspit_b200/, include/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgraspit_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"oracle library missing: {LIB_PATH} (run `make -C oracle` where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_add_body.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.ref_clone_body.argtypes = [C.c_void_p, C.c_int]
        L.ref_remove_body.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_transform.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_activate_body.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_activate_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_is_active.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_all_collisions.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_int]
        L.ref_body_distance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_body_distance.restype = C.c_double
        L.ref_point_distance.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_point_distance.restype = C.c_double
        L.ref_contact.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int]
        L.ref_contact_raw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
        L.ref_fk_chain.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_fk_child_base.argtypes = [C.POINTER(C.c_double)] * 3
        L.ref_state_to_base.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 5
        L.ref_brute_distance.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_brute_distance.restype = C.c_double
        L.ref_contact_trace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.c_int,
                                        C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.c_int]
        L.ref_all_contacts.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.c_int]
        L.ref_region.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double,
                                 C.POINTER(C.c_double), C.c_int]
        L.ref_bounding_volumes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.ref_collide_pair_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ref_distance_pair_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ref_distance_pair_stats.restype = C.c_double
        L.ref_point_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.ref_point_stats.restype = C.c_double
        for name in ("ref_transf_compose",):
            getattr(L, name).argtypes = [C.POINTER(C.c_double)] * 5
        L.ref_transf_inverse.argtypes = [C.POINTER(C.c_double)] * 3
        L.ref_transf_apply.argtypes = [C.POINTER(C.c_double)] * 4
        L.ref_transf_rotate.argtypes = [C.POINTER(C.c_double)] * 4
        L.ref_transf_axis_angle.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_transf_coordinate.argtypes = [C.POINTER(C.c_double)] * 5
        L.ref_transf_jacobian.argtypes = [C.POINTER(C.c_double)] * 3
        L.ref_soft_fit.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                   C.POINTER(C.c_double)]
        L.ref_quat_roundtrip.argtypes = [C.POINTER(C.c_double)] * 3
        L.ref_eval_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                     C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_ubyte),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


class RefWorld:
    """One reference ``GraspitCollision`` instance with integer body ids."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_create())

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- bodies -----------------------------------------------------------------------
    def add_body(self, tris) -> int:
        t = _d(np.asarray(tris, dtype=np.float64))
        assert t.size % 9 == 0
        bid = self.L.ref_add_body(self.h, _dp(t), t.size // 9)
        if bid < 0:
            raise RuntimeError("ref_add_body failed")
        return bid

    def clone_body(self, original: int) -> int:
        return self.L.ref_clone_body(self.h, original)

    def remove_body(self, bid: int):
        self.L.ref_remove_body(self.h, bid)

    def set_transform(self, bid: int, q_wxyz, t):
        q = _d(q_wxyz, 4); tt = _d(t, 3)
        self.L.ref_set_transform(self.h, bid, _dp(q), _dp(tt))

    def activate_body(self, bid, active):
        self.L.ref_activate_body(self.h, bid, int(bool(active)))

    def activate_pair(self, a, b, active):
        self.L.ref_activate_pair(self.h, a, b, int(bool(active)))

    def is_active(self, a, b=-1) -> bool:
        return bool(self.L.ref_is_active(self.h, a, b))

    # --- queries ----------------------------------------------------------------------
    def all_collisions(self, fast=False, interest: Optional[Sequence[int]] = None, cap=4096):
        il = np.ascontiguousarray(np.asarray(interest if interest is not None else [], dtype=np.int32))
        out = np.zeros(2 * cap, dtype=np.int32)
        n = self.L.ref_all_collisions(self.h, int(fast), _ip(il) if interest is not None else None, il.size,
                                      _ip(out), cap)
        pairs = [tuple(sorted((int(out[2 * i]), int(out[2 * i + 1])))) for i in range(0 if fast else min(n, cap))]
        return n, pairs

    def body_distance(self, a, b):
        p1 = np.zeros(3); p2 = np.zeros(3)
        d = self.L.ref_body_distance(self.h, a, b, _dp(p1), _dp(p2))
        return d, p1, p2

    def brute_distance(self, a, b):
        """exhaustive minimum over every triangle pair with the reference's leaf arithmetic (no tree): (dist, mP1 in a's
        frame, mP2 in b's frame, (triangle of a, triangle of b))"""
        p1 = np.zeros(3); p2 = np.zeros(3); tp = np.zeros(2, dtype=np.int32)
        d = self.L.ref_brute_distance(self.h, a, b, _dp(p1), _dp(p2), _ip(tp))
        return d, p1, p2, (int(tp[0]), int(tp[1]))

    def point_distance(self, body, p):
        pp = _d(p, 3); cp = np.zeros(3); cn = np.zeros(3)
        d = self.L.ref_point_distance(self.h, body, _dp(pp), _dp(cp), _dp(cn))
        return d, cp, cn

    def contact(self, a, b, thr, cap=8192):
        out = np.zeros(13 * cap)
        n = self.L.ref_contact(self.h, a, b, float(thr), _dp(out), cap)
        return out[: 13 * min(n, cap)].reshape(-1, 13).copy()

    def contact_raw(self, a, b, thr, cap=65536):
        out = np.zeros(13 * cap)
        st = np.zeros(3, dtype=np.int32)
        n = self.L.ref_contact_raw(self.h, a, b, float(thr), _dp(out), cap, _ip(st))
        return out[: 13 * min(n, cap)].reshape(-1, 13).copy(), st

    def contact_trace(self, a, b, thr, tris_a, tris_b, cap_visits=1 << 16, cap_contacts=1 << 17):
        """leaf pairs in the reference's visit order that appended contacts: (n,3) rows (triangle of a, triangle of b,
        contacts appended) + the raw report in the same order"""
        ta = _d(np.asarray(tris_a, dtype=np.float64)); tb = _d(np.asarray(tris_b, dtype=np.float64))
        vis = np.zeros(3 * cap_visits, dtype=np.int32)
        out = np.zeros(13 * cap_contacts)
        n = self.L.ref_contact_trace(self.h, a, b, float(thr), _dp(ta), ta.size // 9, _dp(tb), tb.size // 9,
                                     _ip(vis), cap_visits, _dp(out), cap_contacts)
        assert n <= cap_visits
        vis = vis[: 3 * n].reshape(-1, 3).copy()
        return vis, out[: 13 * int(vis[:, 2].sum())].reshape(-1, 13).copy()

    def all_contacts(self, thr, interest: Optional[Sequence[int]] = None, cap_pairs=1024, cap_contacts=65536):
        il = np.ascontiguousarray(np.asarray(interest if interest is not None else [], dtype=np.int32))
        pid = np.zeros(2 * cap_pairs, dtype=np.int32); cnt = np.zeros(cap_pairs, dtype=np.int32)
        out = np.zeros(13 * cap_contacts)
        n = self.L.ref_all_contacts(self.h, float(thr), _ip(il) if interest is not None else None, il.size,
                                    _ip(pid), _ip(cnt), cap_pairs, _dp(out), cap_contacts)
        res, k = [], 0
        for i in range(n):
            c = int(cnt[i])
            res.append(((int(pid[2 * i]), int(pid[2 * i + 1])), out[13 * k: 13 * (k + c)].reshape(-1, 13).copy()))
            k += c
        return res

    def region(self, body, p, n, radius, cap=65536):
        pp = _d(p, 3); nn = _d(n, 3); out = np.zeros(3 * cap)
        k = self.L.ref_region(self.h, body, _dp(pp), _dp(nn), float(radius), _dp(out), cap)
        return out[: 3 * min(k, cap)].reshape(-1, 3).copy()

    def bounding_volumes(self, body, depth, cap=1 << 16):
        out = np.zeros(10 * cap)
        k = self.L.ref_bounding_volumes(self.h, body, depth, _dp(out), cap)
        return out[: 10 * min(k, cap)].reshape(-1, 10).copy()

    def collide_pair_stats(self, a, b):
        st = np.zeros(3, dtype=np.int32)
        r = self.L.ref_collide_pair_stats(self.h, a, b, _ip(st))
        return bool(r), st

    def distance_pair_stats(self, a, b):
        st = np.zeros(3, dtype=np.int32)
        d = self.L.ref_distance_pair_stats(self.h, a, b, _ip(st))
        return d, st

    def point_stats(self, body, p):
        st = np.zeros(3, dtype=np.int32); pp = _d(p, 3)
        d = self.L.ref_point_stats(self.h, body, _dp(pp), _ip(st))
        return d, st


# --- transf helpers (pin the shim against test/transform_tests.cpp, test/eigen_tests.cpp) ---

def transf_compose(a, b):
    qa, ta = _d(a[0], 4), _d(a[1], 3); qb, tb = _d(b[0], 4), _d(b[1], 3)
    out = np.zeros(7)
    lib().ref_transf_compose(_dp(qa), _dp(ta), _dp(qb), _dp(tb), _dp(out))
    return out[:4].copy(), out[4:].copy()


def transf_inverse(a):
    q, t = _d(a[0], 4), _d(a[1], 3); out = np.zeros(7)
    lib().ref_transf_inverse(_dp(q), _dp(t), _dp(out))
    return out[:4].copy(), out[4:].copy()


def transf_apply(a, p):
    q, t, pp = _d(a[0], 4), _d(a[1], 3), _d(p, 3); out = np.zeros(3)
    lib().ref_transf_apply(_dp(q), _dp(t), _dp(pp), _dp(out))
    return out


def transf_rotate(a, v):
    q, t, vv = _d(a[0], 4), _d(a[1], 3), _d(v, 3); out = np.zeros(3)
    lib().ref_transf_rotate(_dp(q), _dp(t), _dp(vv), _dp(out))
    return out


def transf_axis_angle(angle, axis):
    ax = _d(axis, 3); out = np.zeros(7)
    lib().ref_transf_axis_angle(float(angle), _dp(ax), _dp(out))
    return out[:4].copy(), out[4:].copy()


def transf_coordinate(o, x, y):
    """the reference's transf::COORDINATE: returns (q wxyz, t, affine 3x3)"""
    o, x, y = _d(o, 3), _d(x, 3), _d(y, 3); out = np.zeros(7); aff = np.zeros(9)
    lib().ref_transf_coordinate(_dp(o), _dp(x), _dp(y), _dp(out), _dp(aff))
    return out[:4].copy(), out[4:].copy(), aff.reshape(3, 3).copy()


def soft_fit(pos, normal, nghbd):
    """SoftContact::SoftContact + FitPoints with the reference's own FitParaboloid / RotateParaboloid (FitParabola.h):
    returns (a, b, c, r1, r2, rotAngle, fitRot 3x3)"""
    p, n = _d(pos, 3), _d(normal, 3)
    pts = np.ascontiguousarray(np.asarray(nghbd, dtype=np.float64).reshape(-1, 3))
    out = np.zeros(15)
    lib().ref_soft_fit(_dp(p), _dp(n), _dp(pts if len(pts) else np.zeros(3)), len(pts), _dp(out))
    return out[:6].copy(), out[6:].reshape(3, 3).copy()


def transf_jacobian(a):
    q, t = _d(a[0], 4), _d(a[1], 3); out = np.zeros(36)
    lib().ref_transf_jacobian(_dp(q), _dp(t), _dp(out))
    return out


def quat_roundtrip(q):
    qq = _d(q, 4); m = np.zeros(9); qo = np.zeros(4)
    lib().ref_quat_roundtrip(_dp(qq), _dp(m), _dp(qo))
    return m.reshape(3, 3), qo


def eval_batch(worlds: Sequence[RefWorld], hand_ids, object_id, tf7, vc_link, vc_loc, vc_normal, mode=0,
               want_dist=False):
    """Reference per-pose loop (legal + CONTACT_ENERGY) over precomputed link poses.

    tf7: (npose, nhand, 7) doubles.  Returns (legal uint8[npose], energy f64[npose], link_dist or None)."""
    L = lib()
    tf7 = np.ascontiguousarray(tf7, dtype=np.float64)
    npose, nhand = tf7.shape[0], tf7.shape[1]
    hid = np.ascontiguousarray(hand_ids, dtype=np.int32)
    assert hid.size == nhand
    vl = np.ascontiguousarray(vc_link, dtype=np.int32)
    vloc = np.ascontiguousarray(vc_loc, dtype=np.float64)
    vn = np.ascontiguousarray(vc_normal, dtype=np.float64)
    legal = np.zeros(npose, dtype=np.uint8)
    energy = np.zeros(npose, dtype=np.float64)
    dist = np.zeros((npose, nhand), dtype=np.float64) if want_dist else None
    arr = (C.c_void_p * len(worlds))(*[w.h for w in worlds])
    L.ref_eval_batch(arr, len(worlds), npose, nhand, _ip(hid), int(object_id), _dp(tf7), vl.size, _ip(vl),
                     _dp(vloc), _dp(vn), int(mode), int(bool(want_dist)),
                     legal.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(energy), _dp(dist) if want_dist else None)
    return legal, energy, dist


# --- FK and state -> pose through the reference's own lines (oracle/ref_fk.cpp) ----------------------------------------

def _t7(tr):
    return np.ascontiguousarray(np.concatenate([np.asarray(tr[0], dtype=np.float64).ravel(), np.asarray(tr[1], dtype=np.float64).ravel()]))


def state_to_base(kind, variables, ref_tran, approach, params=(80.0, 80.0, 160.0)):
    """HandObjectState::execute with the reference's PositionState*::getCoreTran: kind 0 COMPLETE, 1 AXIS_ANGLE, 2 ELLIPSOID,
    3 APPROACH -> (q wxyz, t)"""
    v = _d(variables); pr = _d(params); r7 = _t7(ref_tran); a7 = _t7(approach); out = np.zeros(7)
    rc = lib().ref_state_to_base(int(kind), _dp(v), _dp(pr), _dp(r7), _dp(a7), _dp(out))
    assert rc == 0
    return out[:4].copy(), out[4:].copy()


def forward_kinematics(robots, ridx, base, dofs_by_robot, joint_values=None, out=None):
    """One pose through KinematicChain::fwdKinematics (the reference's lines) for robot ``ridx`` and the robots attached
    to it.  robots: scene RobotDesc list; base = (q, t); dofs_by_robot: {robot: (nDOF,)}; joint_values (optional):
    {robot: per-chain lists of joint values} overriding ratio * dof.  Returns {scene body: (q, t)}."""
    L = lib()
    if out is None:
        out = {}
    r = robots[ridx]
    out[r.base_body] = (np.asarray(base[0], dtype=np.float64).copy(), np.asarray(base[1], dtype=np.float64).copy())
    if getattr(r, "mount_body", -1) >= 0:
        out[r.mount_body] = out[r.base_body]
    owner7 = _t7(base)
    for ci, c in enumerate(r.chains):
        nj, nl = len(c.joints), len(c.link_bodies)
        j6 = np.zeros((max(nj, 1), 6))
        for j, jd in enumerate(c.joints):
            # JointDesc folds the DOF-expression offset c into theta (revolute) / d (prismatic): undo that, the reference keeps them apart
            j6[j] = [1.0, jd.theta - jd.offset, jd.d, jd.a, jd.alpha, jd.offset] if jd.revolute else [0.0, jd.theta, jd.d - jd.offset, jd.a, jd.alpha, jd.offset]
        if joint_values is not None and ridx in joint_values:
            jv = np.ascontiguousarray(joint_values[ridx][ci], dtype=np.float64)
        else:
            d = np.asarray(dofs_by_robot[ridx], dtype=np.float64)
            jv = np.ascontiguousarray([d[jd.dof] * jd.ratio for jd in c.joints], dtype=np.float64)
        lj = np.ascontiguousarray(c.last_joint, dtype=np.int32)
        o7 = np.zeros((nl, 7))
        j6 = np.ascontiguousarray(j6)
        L.ref_fk_chain(_dp(owner7), _dp(_t7(c.tran)), nj, _dp(j6), _dp(jv if nj else np.zeros(1)), nl, _ip(lj), _dp(o7))
        for l, b in enumerate(c.link_bodies):
            out[b] = (o7[l, :4].copy(), o7[l, 4:].copy())
        for (child, off) in getattr(c, "children", []):
            cb = np.zeros(7)
            L.ref_fk_child_base(_dp(np.ascontiguousarray(o7[nl - 1])), _dp(_t7(off)), _dp(cb))
            forward_kinematics(robots, child, (cb[:4].copy(), cb[4:].copy()), dofs_by_robot, joint_values, out)
    return out


def hardware_threads() -> int:
    return int(lib().ref_hardware_threads())


# ---- contact set-up pinned to the reference's own lines (oracle/ref_contact.cpp -> oracle/_ref/libgraspit_ref_contact.so)
CONTACT_LIB_PATH = os.path.join(_HERE, "_ref", "libgraspit_ref_contact.so")
_clib = None


def contact_lib():
    global _clib
    if _clib is None:
        if not os.path.exists(CONTACT_LIB_PATH):
            raise RuntimeError(f"oracle library missing: {CONTACT_LIB_PATH} (run `make -C oracle` where /root/reference exists)")
        L = C.CDLL(CONTACT_LIB_PATH)
        L.ref_soft_neighborhood_radius.restype = C.c_double
        L.ref_soft_neighborhood_radius.argtypes = [C.c_double, C.c_double]
        _clib = L
    return _clib


def _t7(tran):
    return np.ascontiguousarray(np.concatenate([np.asarray(tran[0], dtype=np.float64), np.asarray(tran[1], dtype=np.float64)]))


def point_contact(pos, normal, cof, cog, max_radius):
    """PointContact::PointContact + computeWrenches: (frame q wxyz + t, edges (8, 6), wrenches (8, 6) = force, torque)"""
    frame, edges, wr = np.zeros(7), np.zeros(48), np.zeros(48)
    n = contact_lib().ref_point_contact(_dp(_d(pos, 3)), _dp(_d(normal, 3)), C.c_double(cof), _dp(_d(cog, 3)), C.c_double(max_radius),
                                        _dp(frame), _dp(edges), _dp(wr))
    return frame, edges.reshape(8, 6)[:n], wr.reshape(8, 6)[:n]


def check_contact_normals_ref(row12, tran1, tran2):
    """checkContactNormals (contactSetting.cpp:54-81): returns (row with possibly flipped normals, ok)"""
    r = np.ascontiguousarray(np.asarray(row12, dtype=np.float64)[:12].copy())
    ok = contact_lib().ref_check_contact_normals(_dp(_t7(tran1)), _dp(_t7(tran2)), _dp(r))
    return r, bool(ok)


def soft_neighborhood_radius_ref(y1, y2):
    return float(contact_lib().ref_soft_neighborhood_radius(float(y1), float(y2)))


def merge_soft_neighborhoods_ref(rows12, tran1, tran2, nghbd1, nghbd2):
    """mergeSoftNeighborhoods (contactSetting.cpp:139-186): (surviving original indices, merged nghbd1 list, nghbd2 list)"""
    rows = np.ascontiguousarray(np.asarray(rows12, dtype=np.float64)[:, :12])
    n = len(rows)

    def flat(lists):
        off = np.zeros(n + 1, dtype=np.int32)
        for i, a in enumerate(lists):
            off[i + 1] = off[i] + len(np.asarray(a).reshape(-1, 3))
        x = np.zeros((max(int(off[-1]), 1), 3))
        for i, a in enumerate(lists):
            x[off[i]:off[i + 1]] = np.asarray(a, dtype=np.float64).reshape(-1, 3)
        return np.ascontiguousarray(x), off
    x1, o1 = flat(nghbd1); x2, o2 = flat(nghbd2)
    keep = np.zeros(n, dtype=np.int32)
    out1, out2 = np.zeros((len(x1) + 1, 3)), np.zeros((len(x2) + 1, 3))
    oo1, oo2 = np.zeros(n + 1, dtype=np.int32), np.zeros(n + 1, dtype=np.int32)
    k = contact_lib().ref_merge_soft_neighborhoods(_dp(_t7(tran1)), _dp(_t7(tran2)), n, _dp(rows), _dp(x1), _ip(o1), _dp(x2), _ip(o2),
                                                   _ip(keep), _dp(out1), _ip(oo1), _dp(out2), _ip(oo2))
    return (keep[:k].copy(), [out1[oo1[i]:oo1[i + 1]].copy() for i in range(k)], [out2[oo2[i]:oo2[i + 1]].copy() for i in range(k)])


def soft_contact_pair(tran1, tran2, pos1, nrm1, nghbd1, pos2, nrm2, nghbd2, y1, y2, cof, cog1, max_radius1):
    """SoftContact + mate as addContacts builds them, setUpFrictionEdges + computeWrenches on side 1:
    (wrenches (20, 6), info side 1, info side 2); info = majorAxis, minorAxis, relPhi, r1prime, r2prime, a, b, c, r1, r2"""
    a = np.ascontiguousarray(np.asarray(nghbd1, dtype=np.float64).reshape(-1, 3)); b = np.ascontiguousarray(np.asarray(nghbd2, dtype=np.float64).reshape(-1, 3))
    wr, i1, i2 = np.zeros(20 * 6), np.zeros(10), np.zeros(10)
    n = contact_lib().ref_soft_contact_pair(_dp(_t7(tran1)), _dp(_t7(tran2)), _dp(_d(pos1, 3)), _dp(_d(nrm1, 3)), len(a), _dp(a if len(a) else np.zeros(3)),
                                            _dp(_d(pos2, 3)), _dp(_d(nrm2, 3)), len(b), _dp(b if len(b) else np.zeros(3)),
                                            C.c_double(y1), C.c_double(y2), C.c_double(cof), _dp(_d(cog1, 3)), C.c_double(max_radius1),
                                            _dp(wr), _dp(i1), _dp(i2))
    return wr.reshape(20, 6)[:n], i1, i2


# ---- simulated annealing pinned to the reference's own SimAnn (oracle/ref_simann.cpp -> libgraspit_ref_simann.so)
SIMANN_LIB_PATH = os.path.join(_HERE, "_ref", "libgraspit_ref_simann.so")
_slib = None


def simann_run(neg, seed, preset, iters, start, w, c, cut, start_energy):
    """`iters` calls of the reference's SimAnn::iterate on one SPACE_AXIS_ANGLE + POSE_EIGEN state after srand(seed), with
    the closed-form energy of oracle/ref_simann.cpp.  Variables position first, then posture.  Returns dict(result
    (0 FAIL, 1 JUMP, 2 KEEP), state (iters, nvar), energy, step, table (nvar, 4) = min, max, max jump, circular)."""
    global _slib
    if _slib is None:
        if not os.path.exists(SIMANN_LIB_PATH):
            raise RuntimeError(f"oracle library missing: {SIMANN_LIB_PATH} (run `make -C oracle` where /root/reference exists)")
        _slib = C.CDLL(SIMANN_LIB_PATH)
    nvar = 6 + neg
    res = np.zeros(iters, dtype=np.int32); st = np.zeros((iters, nvar)); en = np.zeros(iters); stp = np.zeros(iters, dtype=np.int32)
    table = np.zeros((nvar, 4))
    rc = _slib.ref_simann_run(int(neg), C.c_uint(seed), int(preset), int(iters), _dp(_d(start, nvar)), _dp(_d(w, nvar)), _dp(_d(c, nvar)),
                              C.c_double(cut), C.c_double(start_energy), _ip(res), _dp(st), _dp(en), _ip(stp), _dp(table))
    assert rc == nvar, rc
    return dict(result=res, state=st, energy=en, step=stp, table=table)


# ---- autograsp pinned to the reference's own lines (oracle/ref_autograsp.cpp, inside libgraspit_ref.so) ------------------
class RefAutoGrasp:
    """One hand (no attached robots) and the rest of a scene inside the reference's collision backend, closed by the
    reference's own Hand::autoGrasp / Robot::moveDOFToContacts / DOF::accumulateMove lines."""

    DOF_TYPES = {"r": 0, "b": 1, "c": 2}

    def __init__(self, scene, ridx=0, dof_type=None, stiffness=1.0):
        L = lib()
        L.ref_ag_create.restype = C.c_void_p
        self.L, self.scene, self.r = L, scene, scene.robots[ridx]
        r = self.r
        assert not any(c.children for c in r.chains), "RefAutoGrasp: single-robot trees only"
        self.h = C.c_void_p(L.ref_ag_create())
        self.ids = {}
        for i, b in enumerate(scene.bodies):
            if b.robot == ridx and b.kind == "link":
                chain, link = b.chain, b.link
            elif b.robot == ridx:
                chain, link = -1, 0                     # base / mount piece
            else:
                chain, link = -2, 0
            tris = np.ascontiguousarray(b.tris.astype(np.float64).reshape(-1, 9))
            self.ids[i] = L.ref_ag_add_body(self.h, _dp(tris), len(tris), chain, link)
            L.ref_ag_set_transform(self.h, self.ids[i], _dp(_t7(scene.body_tran[i])))
        for (a, b) in list(scene.disabled_pairs) + list(scene.touching_pairs):
            L.ref_ag_activate_pair(self.h, self.ids[a], self.ids[b], 0)
        nch = len(r.chains)
        ctran = np.ascontiguousarray(np.stack([_t7(c.tran) for c in r.chains]))
        cnj = np.ascontiguousarray([len(c.joints) for c in r.chains], dtype=np.int32)
        cnl = np.ascontiguousarray([len(c.link_bodies) for c in r.chains], dtype=np.int32)
        rows = []
        for c in r.chains:
            for jd in c.joints:
                dh = [1.0, jd.theta - jd.offset, jd.d, jd.a, jd.alpha, jd.offset] if jd.revolute else [0.0, jd.theta, jd.d - jd.offset, jd.a, jd.alpha, jd.offset]
                coupling = jd.coupling if getattr(jd, "coupling", None) is not None else jd.ratio
                k = jd.stiffness if getattr(jd, "stiffness", 0.0) else stiffness
                rows.append([jd.dof, coupling] + dh + [jd.min, jd.max, k])
        joints = np.ascontiguousarray(np.array(rows, dtype=np.float64))
        llj = np.ascontiguousarray([j for c in r.chains for j in c.last_joint], dtype=np.int32)
        lbody = np.ascontiguousarray([self.ids[b] for c in r.chains for b in c.link_bodies], dtype=np.int32)
        types = dof_type if dof_type is not None else r.dof_type
        dtype_codes = np.ascontiguousarray([self.DOF_TYPES[t] for t in types], dtype=np.int32)
        vel = np.ascontiguousarray(np.asarray(r.dof_velocity, dtype=np.float64))
        self.dof_min, self.dof_max = np.zeros(r.n_dof), np.zeros(r.n_dof)
        self.nj = L.ref_ag_define_hand(self.h, nch, _dp(ctran), _ip(cnj), _ip(cnl), _dp(joints), _ip(llj), _ip(lbody), r.n_dof, _ip(dtype_codes),
                                       _dp(vel), self.ids[r.base_body], self.ids[r.mount_body] if r.mount_body >= 0 else -1,
                                       _dp(self.dof_min), _dp(self.dof_max))

    def auto_grasp(self, base, dof0, speed_factor=1.0, stop_at_contact=False):
        """Hand::autoGrasp(false, speed_factor, stop_at_contact) from (base = (q, t), dof0): dict(dof, joints, stopped, moved,
        jumps (successful jumpDOFToContact calls), error ("Interpolation failed!"), failsafe)"""
        dof, jv, st, info = np.zeros(self.r.n_dof), np.zeros(self.nj), np.zeros(self.nj, dtype=np.int32), np.zeros(4, dtype=np.int32)
        self.L.ref_ag_autograsp(self.h, _dp(_t7(base)), _dp(_d(dof0, self.r.n_dof)), C.c_double(speed_factor), int(stop_at_contact),
                                _dp(dof), _dp(jv), _ip(st), _ip(info))
        return dict(dof=dof, joints=jv, stopped=st, moved=bool(info[0]), jumps=int(info[1]), error=bool(info[2]), failsafe=bool(info[3]))

    def close(self):
        if self.h:
            self.L.ref_ag_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
