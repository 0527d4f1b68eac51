"""Python host mirror of the reference's ``CollisionInterface`` over the C ABI (``include/b2c.h``).

Method names and argument meaning follow ``include/graspit/Collision/collisionInterface.h:44-185``
(``addBody``, ``setBodyTransform``, ``activatePair``, ``allCollisions``, ``bodyToBodyDistance``,
``pointToBodyDistance``, ``contact``, ``allContacts``, ``bodyRegion``, ``getBoundingVolumes``) so that
the parity tests read like the reference's own ``test/graspit_collision_tests.cpp``; the batched
``eval_batch`` replaces the ``SearchEnergy::analyzeState`` loop.  Everything here is plumbing: the
work happens in ``libb2c.so`` (CUDA, sm_100a).  There is no CPU path -- if the library or a CUDA
device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2C_LIB selects a tuning variant of the same engine (graspit_b200/build.py --out=...); default = the product build
LIB_PATH = os.environ.get("B2C_LIB") or os.path.join(_HERE, "libb2c.so")

B2C_INPUT_RAW, B2C_INPUT_AA_EIGEN, B2C_INPUT_JOINTS = 0, 1, 2
B2C_INPUT_ELLIPSOID_EIGEN, B2C_INPUT_APPROACH_EIGEN, B2C_INPUT_COMPLETE_DOF = 3, 4, 5
B2C_EVAL_PRESET, B2C_EVAL_LIVE, B2C_EVAL_PAIRMASK, B2C_EVAL_LINKDIST, B2C_EVAL_DEVICE_PTRS = 0, 1, 2, 4, 8
FAST_COLLISION, ALL_COLLISIONS = 0, 1

EXPORTS = [
    "b2c_create", "b2c_destroy", "b2c_last_error", "b2c_stream", "b2c_device", "b2c_launch_count",
    "b2c_set_stream", "b2c_set_profiling", "b2c_last_kernel_ms",
    "b2c_mesh_create", "b2c_mesh_destroy", "b2c_body_create", "b2c_body_set_mesh", "b2c_body_remove",
    "b2c_body_set_transform", "b2c_body_get_transform", "b2c_body_set_active", "b2c_pair_set_active",
    "b2c_pair_is_active", "b2c_body_set_thread", "b2c_set_thread", "b2c_all_collisions", "b2c_body_distance", "b2c_point_distance", "b2c_contacts",
    "b2c_all_contacts", "b2c_region", "b2c_bounding_volumes", "b2c_robot_define", "b2c_eval_batch",
    "b2c_robot_bodies", "b2c_robot_pairs", "b2c_last_stats", "b2c_obb_hierarchy", "b2c_contacts_postprocess",
    "b2c_contact_visit_order", "b2c_check_overflow",
    "b2c_exchange_create", "b2c_exchange_connect", "b2c_anneal_publish", "b2c_exchange_collect", "b2c_exchange_destroy",
    "b2c_anneal_create", "b2c_anneal_destroy", "b2c_anneal_step", "b2c_anneal_read", "b2c_anneal_device_ptrs",
    "b2c_anneal_propose", "b2c_anneal_write", "b2c_anneal_best_rows",
    "b2c_check_contact_normals", "b2c_contact_frames", "b2c_point_contact_wrenches",
    "b2c_autograsp_batch", "b2c_contacts_batch", "b2c_region_batch",
    "b2c_soft_neighborhood_radius", "b2c_soft_contact_merge", "b2c_soft_contact_fit", "b2c_soft_contact_wrenches",
]


class B2CError(RuntimeError):
    pass


class Contact(C.Structure):
    _fields_ = [("b1_pos", C.c_double * 3), ("b2_pos", C.c_double * 3), ("b1_normal", C.c_double * 3),
                ("b2_normal", C.c_double * 3), ("distSq", C.c_double)]


class SoftFit(C.Structure):
    _fields_ = [("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("r1", C.c_double), ("r2", C.c_double),
                ("rot_angle", C.c_double)]


class Obb(C.Structure):
    _fields_ = [("t", C.c_double * 3), ("q_wxyz", C.c_double * 4), ("half", C.c_double * 3)]


class JointDesc(C.Structure):
    _fields_ = [("dof", C.c_int), ("revolute", C.c_int), ("ratio", C.c_double), ("theta", C.c_double),
                ("d", C.c_double), ("a", C.c_double), ("alpha", C.c_double)]


class ChainDesc(C.Structure):
    _fields_ = [("tran_q", C.c_double * 4), ("tran_t", C.c_double * 3), ("first_joint", C.c_int),
                ("njoints", C.c_int), ("first_link", C.c_int), ("nlinks", C.c_int)]


class VContactDesc(C.Structure):
    _fields_ = [("body", C.c_int), ("loc", C.c_double * 3), ("normal", C.c_double * 3)]


class RobotDescC(C.Structure):
    _fields_ = [("base_body", C.c_int), ("mount_body", C.c_int), ("ndof", C.c_int),
                ("dof_min", C.POINTER(C.c_double)), ("dof_max", C.POINTER(C.c_double)),
                ("nchains", C.c_int), ("chains", C.POINTER(ChainDesc)),
                ("njoints", C.c_int), ("joints", C.POINTER(JointDesc)),
                ("nlinks", C.c_int), ("link_body", C.POINTER(C.c_int)), ("link_last_joint", C.POINTER(C.c_int)),
                ("parent_robot", C.c_int), ("parent_chain", C.c_int),
                ("parent_offset_q", C.c_double * 4), ("parent_offset_t", C.c_double * 3),
                ("approach_q", C.c_double * 4), ("approach_t", C.c_double * 3),
                ("neg", C.c_int), ("eg", C.POINTER(C.c_double)),
                ("eg_origin", C.POINTER(C.c_double)), ("eg_norm", C.POINTER(C.c_double)),
                ("nvcontacts", C.c_int), ("vcontacts", C.POINTER(VContactDesc))]


class EvalArgs(C.Structure):
    _fields_ = [("robot", C.c_int), ("object", C.c_int), ("npose", C.c_int), ("input_mode", C.c_int),
                ("states", C.c_void_p), ("stride", C.c_int),
                ("ref_q", C.c_double * 4), ("ref_t", C.c_double * 3), ("flags", C.c_int),
                ("legal", C.c_void_p), ("energy", C.c_void_p), ("pair_mask", C.c_void_p),
                ("link_dist", C.c_void_p), ("body_tf", C.c_void_p), ("position_params", C.c_double * 3)]


class ContactsBatchArgs(C.Structure):
    _fields_ = [("robot", C.c_int), ("npose", C.c_int), ("input_mode", C.c_int), ("states", C.c_void_p), ("stride", C.c_int),
                ("ref_q", C.c_double * 4), ("ref_t", C.c_double * 3), ("threshold", C.c_double), ("raw", C.c_int),
                ("contacts", C.c_void_p), ("contact_pose", C.c_void_p), ("contact_pair", C.c_void_p), ("cap_contacts", C.c_int),
                ("ncontacts", C.c_void_p), ("pose_offsets", C.c_void_p)]


class AutoGraspArgs(C.Structure):
    _fields_ = [("robot", C.c_int), ("npose", C.c_int), ("states", C.c_void_p), ("stride", C.c_int),
                ("desired", C.c_void_p), ("steps", C.c_void_p), ("dof_breakaway", C.c_void_p),
                ("stop_at_contact", C.c_int), ("contact_threshold", C.c_double), ("max_rounds", C.c_int),
                ("dof_out", C.c_void_p), ("joint_out", C.c_void_p), ("status", C.c_void_p), ("stopped_out", C.c_void_p),
                ("rounds", C.c_void_p)]


AG_MOVED, AG_INTERP_ERROR, AG_FAILSAFE, AG_STOPPED_AT_CONTACT = 1, 2, 4, 8
AUTO_GRASP_TIME_STEP = 0.01       # Robot::AUTO_GRASP_TIME_STEP (reference src/robot.cpp:68)


class AnnealDesc(C.Structure):
    _fields_ = [("robot", C.c_int), ("object", C.c_int), ("nchain", C.c_int), ("nvar", C.c_int),
                ("vmin", C.POINTER(C.c_double)), ("vmax", C.POINTER(C.c_double)), ("vjump", C.POINTER(C.c_double)),
                ("circular", C.POINTER(C.c_ubyte)), ("ref_q", C.c_double * 4), ("ref_t", C.c_double * 3),
                ("YC", C.c_double), ("HC", C.c_double), ("NBR_ADJ", C.c_double), ("ERR_ADJ", C.c_double), ("T0", C.c_double),
                ("YDIMS", C.c_int), ("HDIMS", C.c_int), ("K0", C.c_int), ("seed", C.c_ulonglong),
                ("chain_offset", C.c_int), ("flags", C.c_int)]


_lib = None


def load_library():
    """dlopen libb2c.so and declare prototypes.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2CError(f"CUDA extension missing: {LIB_PATH}. Build it with `python -m graspit_b200.build` "
                       "(there is no CPU fallback).")
    L = C.CDLL(LIB_PATH)
    vp, ci, dp, ip = C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.b2c_create.argtypes = [ci, C.POINTER(vp)]
    L.b2c_destroy.argtypes = [vp]; L.b2c_destroy.restype = None
    L.b2c_last_error.argtypes = [vp]; L.b2c_last_error.restype = C.c_char_p
    L.b2c_stream.argtypes = [vp]; L.b2c_stream.restype = vp
    L.b2c_device.argtypes = [vp]
    L.b2c_launch_count.argtypes = [vp]; L.b2c_launch_count.restype = C.c_longlong
    L.b2c_set_stream.argtypes = [vp, vp]
    L.b2c_set_profiling.argtypes = [vp, ci]
    L.b2c_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.b2c_mesh_create.argtypes = [vp, C.POINTER(C.c_float), ci, ip]
    L.b2c_mesh_destroy.argtypes = [vp, ci]
    L.b2c_body_create.argtypes = [vp, ci, ip]
    L.b2c_body_set_mesh.argtypes = [vp, ci, ci]
    L.b2c_body_remove.argtypes = [vp, ci]
    L.b2c_body_set_transform.argtypes = [vp, ci, dp, dp]
    L.b2c_body_get_transform.argtypes = [vp, ci, dp, dp]
    L.b2c_body_set_active.argtypes = [vp, ci, ci]
    L.b2c_pair_set_active.argtypes = [vp, ci, ci, ci]
    L.b2c_pair_is_active.argtypes = [vp, ci, ci, ip]
    L.b2c_check_contact_normals.argtypes = [C.POINTER(Contact), ci, dp, dp, dp, dp, ip]
    L.b2c_contact_frames.argtypes = [C.POINTER(Contact), ci, ci, dp]
    L.b2c_point_contact_wrenches.argtypes = [C.POINTER(Contact), ci, ci, C.c_double, dp, C.c_double, dp]
    L.b2c_soft_neighborhood_radius.argtypes = [C.c_double, C.c_double]
    L.b2c_soft_neighborhood_radius.restype = C.c_double
    L.b2c_soft_contact_merge.argtypes = [C.POINTER(Contact), ip, dp, dp, dp, dp, dp, ip, dp, ip, ip]
    L.b2c_soft_contact_fit.argtypes = [C.POINTER(Contact), ci, ci, dp, ip, C.POINTER(SoftFit)]
    L.b2c_soft_contact_wrenches.argtypes = [C.POINTER(Contact), ci, ci, C.POINTER(SoftFit), C.POINTER(SoftFit), dp, dp,
                                            C.c_double, C.c_double, dp, C.c_double, dp, dp]
    L.b2c_anneal_create.argtypes = [vp, C.POINTER(AnnealDesc), C.POINTER(C.c_float), ip]
    L.b2c_anneal_destroy.argtypes = [vp, ci]
    L.b2c_anneal_step.argtypes = [vp, ci, ci]
    L.b2c_anneal_read.argtypes = [vp, ci, vp, vp, vp, vp, vp, C.POINTER(C.c_longlong)]
    L.b2c_anneal_device_ptrs.argtypes = [vp, ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.b2c_anneal_best_rows.argtypes = [vp, ci, ci, vp]
    L.b2c_anneal_propose.argtypes = [vp, ci, vp]
    L.b2c_anneal_write.argtypes = [vp, ci, vp, vp, vp, ci]
    L.b2c_body_set_thread.argtypes = [vp, ci, ci]
    L.b2c_set_thread.argtypes = [vp, ci]
    L.b2c_all_collisions.argtypes = [vp, ci, ip, ci, ip, ci, ip]
    L.b2c_body_distance.argtypes = [vp, ci, ci, dp, dp, dp, dp, dp]
    L.b2c_point_distance.argtypes = [vp, ci, dp, dp, dp, dp]
    L.b2c_contacts.argtypes = [vp, ci, ci, C.c_double, ci, C.POINTER(Contact), ci, ip]
    L.b2c_all_contacts.argtypes = [vp, C.c_double, ip, ci, ip, ip, ci, ip, C.POINTER(Contact), ci, ip]
    L.b2c_region.argtypes = [vp, ci, dp, dp, C.c_double, dp, ci, ip]
    L.b2c_bounding_volumes.argtypes = [vp, ci, ci, C.POINTER(Obb), ci, ip]
    L.b2c_robot_define.argtypes = [vp, C.POINTER(RobotDescC), ip]
    L.b2c_eval_batch.argtypes = [vp, C.POINTER(EvalArgs)]
    L.b2c_autograsp_batch.argtypes = [vp, C.POINTER(AutoGraspArgs)]
    L.b2c_contacts_batch.argtypes = [vp, C.POINTER(ContactsBatchArgs)]
    L.b2c_region_batch.argtypes = [vp, ci, ip, dp, dp, dp, dp, ci, ip, ip]
    L.b2c_robot_bodies.argtypes = [vp, ci, ip, ci, ip]
    L.b2c_robot_pairs.argtypes = [vp, ci, ip, ci, ip]
    L.b2c_last_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.b2c_obb_hierarchy.argtypes = [C.POINTER(C.c_float), ci, ci, C.POINTER(Obb), ci, ip]
    L.b2c_contacts_postprocess.argtypes = [C.POINTER(Contact), ip, C.c_double, ci]
    L.b2c_check_overflow.argtypes = [vp]
    L.b2c_exchange_create.argtypes = [vp, ci, ci, ci, ci, ci, vp, C.POINTER(vp)]
    L.b2c_exchange_connect.argtypes = [vp, vp, C.POINTER(vp), ip]
    L.b2c_anneal_publish.argtypes = [vp, ci, vp]
    L.b2c_exchange_collect.argtypes = [vp, vp, vp, ci]
    L.b2c_exchange_destroy.argtypes = [vp]
    L.b2c_contact_visit_order.argtypes = [C.POINTER(C.c_float), ci, C.POINTER(C.c_float), ci, dp, dp, dp, dp, ip, ip, ci, ip]
    _lib = L
    return L


def _d(x, n=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _contacts_to_array(buf, n):
    """ctypes Contact array -> (n, 13) float64 (b2c_contact is 13 consecutive doubles)"""
    if n <= 0:
        return np.zeros((0, 13))
    return np.frombuffer(buf, dtype=np.float64, count=13 * n).reshape(n, 13).copy()


def obb_hierarchy(tris, depth: int, cap=1 << 16):
    """Host-only: the reference-style oriented-box hierarchy of a triangle soup at ``depth`` -> (n,10) rows
    (t, q wxyz, half).  What getBoundingVolumes serves."""
    L = load_library()
    t = np.ascontiguousarray(np.asarray(tris, dtype=np.float32).reshape(-1, 9))
    buf = (Obb * cap)()
    n = C.c_int(0)
    rc = L.b2c_obb_hierarchy(t.ctypes.data_as(C.POINTER(C.c_float)), t.shape[0], depth, buf, cap, C.byref(n))
    if rc != 0:
        raise B2CError(f"b2c_obb_hierarchy failed: {rc}")
    out = np.zeros((min(n.value, cap), 10))
    for i in range(out.shape[0]):
        out[i, 0:3] = buf[i].t[:]; out[i, 3:7] = buf[i].q_wxyz[:]; out[i, 7:10] = buf[i].half[:]
    return out


def contact_visit_order(tris_a, tris_b, tf_a, tf_b, ta, tb):
    """Host-only: the order in which the reference's ContactCallback reaches the triangle pairs (ta[i], tb[i]) of two
    models with world transforms tf_a / tf_b = (q wxyz, t) -> permutation of range(len(ta)) (b2c_contact_visit_order)."""
    L = load_library()
    A = np.ascontiguousarray(np.asarray(tris_a, dtype=np.float32).reshape(-1, 9))
    B = np.ascontiguousarray(np.asarray(tris_b, dtype=np.float32).reshape(-1, 9))
    ia = np.ascontiguousarray(ta, dtype=np.int32); ib = np.ascontiguousarray(tb, dtype=np.int32)
    assert ia.shape == ib.shape
    order = np.zeros(max(1, ia.size), dtype=np.int32)
    qa = (C.c_double * 4)(*[float(v) for v in tf_a[0]]); pa = (C.c_double * 3)(*[float(v) for v in tf_a[1]])
    qb = (C.c_double * 4)(*[float(v) for v in tf_b[0]]); pb = (C.c_double * 3)(*[float(v) for v in tf_b[1]])
    fp = C.POINTER(C.c_float); ip = C.POINTER(C.c_int)
    rc = L.b2c_contact_visit_order(A.ctypes.data_as(fp), A.shape[0], B.ctypes.data_as(fp), B.shape[0], qa, pa, qb, pb,
                                   ia.ctypes.data_as(ip), ib.ctypes.data_as(ip), ia.size, order.ctypes.data_as(ip))
    if rc != 0:
        raise B2CError(f"b2c_contact_visit_order failed: {rc}")
    return order[: ia.size].copy()


def contacts_postprocess(contacts, duplicate_threshold=3.0, compact=True):
    """Host-only: removeContactDuplicates + compactContactSet on an (n,13) contact array."""
    L = load_library()
    c = np.asarray(contacts, dtype=np.float64).reshape(-1, 13)
    n = C.c_int(c.shape[0])
    buf = (Contact * max(1, c.shape[0]))()
    for i in range(c.shape[0]):
        buf[i].b1_pos[:] = list(c[i, 0:3]); buf[i].b2_pos[:] = list(c[i, 3:6])
        buf[i].b1_normal[:] = list(c[i, 6:9]); buf[i].b2_normal[:] = list(c[i, 9:12]); buf[i].distSq = c[i, 12]
    rc = L.b2c_contacts_postprocess(buf, C.byref(n), float(duplicate_threshold), int(compact))
    if rc != 0:
        raise B2CError(f"b2c_contacts_postprocess failed: {rc}")
    return _contacts_to_array(buf, n.value)


class Engine:
    """One collision context on one CUDA device (one instance per World, like the reference)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.b2c_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise B2CError(f"b2c_create(device={device}) failed with code {rc}: no usable CUDA device "
                           "(the engine has no CPU path)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.b2c_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, soft=False):
        if rc != 0 and not soft:
            raise B2CError(f"b2c error {rc}: {self.L.b2c_last_error(self.h).decode()}")
        return rc

    @property
    def launch_count(self) -> int:
        return int(self.L.b2c_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.b2c_stream(self.h) or 0)

    def set_stream(self, cuda_stream: int):
        """launch on a caller-owned CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``)"""
        self._ck(self.L.b2c_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_profiling(self, on: bool):
        self._ck(self.L.b2c_set_profiling(self.h, int(bool(on))))

    def last_kernel_ms(self) -> dict:
        ms = (C.c_float * 8)()
        self._ck(self.L.b2c_last_kernel_ms(self.h, ms))
        return {"fk": ms[0], "eval": ms[1], "recheck": ms[2], "dist": ms[3], "energy": ms[4]}

    # ---- addBody / cloneBody / removeBody / updateBodyGeometry ---------------------------------
    def create_mesh(self, tris) -> int:
        t = np.ascontiguousarray(np.asarray(tris, dtype=np.float32).reshape(-1, 9))
        mid = C.c_int(-1)
        self._ck(self.L.b2c_mesh_create(self.h, t.ctypes.data_as(C.POINTER(C.c_float)), t.shape[0], C.byref(mid)))
        return mid.value

    def add_body(self, tris=None, mesh: Optional[int] = None) -> int:
        """addBody: register geometry (or reuse ``mesh``) and return the body id."""
        if mesh is None:
            mesh = self.create_mesh(tris)
        bid = C.c_int(-1)
        self._ck(self.L.b2c_body_create(self.h, mesh, C.byref(bid)))
        self._mesh_of = getattr(self, "_mesh_of", {})
        self._mesh_of[bid.value] = mesh
        return bid.value

    def clone_body(self, original: int) -> int:
        """cloneBody: a second body sharing the original's collision geometry."""
        return self.add_body(mesh=self._mesh_of[original])

    def update_body_geometry(self, body: int, tris):
        mesh = self.create_mesh(tris)
        self._ck(self.L.b2c_body_set_mesh(self.h, body, mesh))
        self._mesh_of[body] = mesh

    def remove_body(self, body: int):
        self._ck(self.L.b2c_body_remove(self.h, body))

    # ---- setBodyTransform / activation -----------------------------------------------------------
    def set_body_transform(self, body: int, q_wxyz, t):
        q, tt = _d(q_wxyz, 4), _d(t, 3)
        self._ck(self.L.b2c_body_set_transform(self.h, body, _dp(q), _dp(tt)))

    def get_body_transform(self, body: int):
        q, t = np.zeros(4), np.zeros(3)
        self._ck(self.L.b2c_body_get_transform(self.h, body, _dp(q), _dp(t)))
        return q, t

    def activate_body(self, body: int, active: bool):
        self._ck(self.L.b2c_body_set_active(self.h, body, int(bool(active))))

    def activate_pair(self, a: int, b: int, active: bool):
        self._ck(self.L.b2c_pair_set_active(self.h, a, b, int(bool(active))))

    def set_body_thread(self, body: int, thread_id: int):
        """thread that owns the body (CollisionModel thread id; 0 = master)"""
        self._ck(self.L.b2c_body_set_thread(self.h, body, thread_id))

    def set_thread(self, thread_id: int):
        """thread on whose behalf the following queries run (CollisionInterface::getThreadId)"""
        self._ck(self.L.b2c_set_thread(self.h, thread_id))

    def is_active(self, a: int, b: int = -1) -> bool:
        out = C.c_int(0)
        rc = self.L.b2c_pair_is_active(self.h, a, b, C.byref(out))
        return bool(out.value) if rc == 0 else False

    # ---- queries ---------------------------------------------------------------------------------
    def all_collisions(self, detection=ALL_COLLISIONS, interest: Optional[Sequence[int]] = None, cap=4096):
        """allCollisions -> (count, [(a,b), ...]); the report is empty for FAST_COLLISION."""
        il = np.ascontiguousarray(np.asarray(interest if interest is not None else [], dtype=np.int32))
        out = np.zeros(2 * cap, dtype=np.int32)
        n = C.c_int(0)
        self._ck(self.L.b2c_all_collisions(self.h, int(detection == FAST_COLLISION),
                                           _ip(il) if interest is not None else None, il.size, _ip(out), cap, C.byref(n)))
        pairs = [] if detection == FAST_COLLISION else [
            (int(out[2 * i]), int(out[2 * i + 1])) for i in range(min(n.value, cap))]
        return n.value, pairs

    def body_to_body_distance(self, a: int, b: int, local=False):
        """bodyToBodyDistance -> (dist, p1, p2) in the reference's output convention
        (``local=True``: plain closest points in each body's frame)."""
        d = C.c_double(0.0)
        p1, p2, l1, l2 = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
        rc = self.L.b2c_body_distance(self.h, a, b, C.byref(d), _dp(p1), _dp(p2), _dp(l1), _dp(l2))
        if rc == -1:
            return 0.0, p1, p2      # reference: unknown body -> returns 0
        self._ck(rc)
        return (d.value, l1, l2) if local else (d.value, p1, p2)

    def point_to_body_distance(self, body: int, point):
        """pointToBodyDistance -> (dist, closestPoint, closestNormal), world frame."""
        d = C.c_double(0.0)
        p = _d(point, 3); cp, cn = np.zeros(3), np.zeros(3)
        rc = self.L.b2c_point_distance(self.h, body, _dp(p), C.byref(d), _dp(cp), _dp(cn))
        if rc == -1:
            return 0.0, cp, cn
        self._ck(rc)
        return d.value, cp, cn

    def contact(self, a: int, b: int, threshold: float, raw=False, cap=65536):
        buf = (Contact * cap)()
        n = C.c_int(0)
        self._ck(self.L.b2c_contacts(self.h, a, b, float(threshold), int(raw), buf, cap, C.byref(n)))
        return _contacts_to_array(buf, min(n.value, cap))

    def all_contacts(self, threshold: float, interest: Optional[Sequence[int]] = None, cap_pairs=1024,
                     cap_contacts=65536):
        il = np.ascontiguousarray(np.asarray(interest if interest is not None else [], dtype=np.int32))
        pid = np.zeros(2 * cap_pairs, dtype=np.int32); cnt = np.zeros(cap_pairs, dtype=np.int32)
        buf = (Contact * cap_contacts)()
        npairs, nc = C.c_int(0), C.c_int(0)
        self._ck(self.L.b2c_all_contacts(self.h, float(threshold), _ip(il) if interest is not None else None,
                                         il.size, _ip(pid), _ip(cnt), cap_pairs, C.byref(npairs), buf, cap_contacts,
                                         C.byref(nc)))
        allc = _contacts_to_array(buf, min(nc.value, cap_contacts))
        res, k = [], 0
        for i in range(min(npairs.value, cap_pairs)):
            c = int(cnt[i])
            res.append(((int(pid[2 * i]), int(pid[2 * i + 1])), allc[k:k + c]))
            k += c
        return res

    def body_region(self, body: int, point, normal, radius: float, cap=1 << 18):
        p, nn = _d(point, 3), _d(normal, 3)
        out = np.zeros(3 * cap)
        n = C.c_int(0)
        self._ck(self.L.b2c_region(self.h, body, _dp(p), _dp(nn), float(radius), _dp(out), cap, C.byref(n)))
        return out[: 3 * min(n.value, cap)].reshape(-1, 3).copy()

    def body_region_batch(self, bodies, points, normals, radii, cap=None, flat=False):
        """bodyRegion for n (body, point, normal, radius) queries in one call: list of (k_i, 3) arrays"""
        b = np.ascontiguousarray(bodies, dtype=np.int32)
        n = b.size
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(n, 3)
        nn = np.ascontiguousarray(normals, dtype=np.float64).reshape(n, 3)
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, dtype=np.float64), (n,)))
        cap = int(cap) if cap is not None else max(4096, 96 * n)
        out = np.empty((cap, 3)); off = np.zeros(n + 1, dtype=np.int32)
        tot = C.c_int(0)
        self._ck(self.L.b2c_region_batch(self.h, n, _ip(b), _dp(p), _dp(nn), _dp(r), _dp(out), cap, _ip(off), C.byref(tot)))
        if tot.value > cap:
            return self.body_region_batch(bodies, points, normals, radii, cap=tot.value, flat=flat)
        if flat:
            return out[:tot.value], off                 # ragged form of the C ABI: points back to back + offsets[n + 1]
        return [out[off[i]:off[i + 1]].copy() for i in range(n)]

    def get_bounding_volumes(self, body: int, depth: int, cap=1 << 16):
        buf = (Obb * cap)()
        n = C.c_int(0)
        self._ck(self.L.b2c_bounding_volumes(self.h, body, depth, buf, cap, C.byref(n)))
        out = np.zeros((min(n.value, cap), 10))
        for i in range(out.shape[0]):
            out[i, 0:3] = buf[i].t[:]; out[i, 3:7] = buf[i].q_wxyz[:]; out[i, 7:10] = buf[i].half[:]
        return out

    # ---- robots and batched evaluation ----------------------------------------------------------
    def define_robot(self, robot, body_ids: dict, parent_robot_id: int = -1, parent_offset=None) -> int:
        """Register a robot description (``formats.xmlmodel.RobotDesc``).  ``body_ids`` maps scene body
        indices to engine body ids."""
        joints, chains, link_body, link_last = [], [], [], []
        for c in robot.chains:
            cd = ChainDesc()
            cd.tran_q[:] = list(c.tran[0]); cd.tran_t[:] = list(c.tran[1])
            cd.first_joint, cd.njoints = len(joints), len(c.joints)
            cd.first_link, cd.nlinks = len(link_body), len(c.link_bodies)
            for j in c.joints:
                jd = JointDesc(j.dof, int(j.revolute), j.ratio, j.theta, j.d, j.a, j.alpha)
                joints.append(jd)
            link_body += [body_ids[b] for b in c.link_bodies]
            link_last += list(c.last_joint)
            chains.append(cd)
        d = RobotDescC()
        keep = []
        d.base_body = body_ids[robot.base_body]
        d.mount_body = body_ids[robot.mount_body] if robot.mount_body >= 0 else -1
        d.ndof = robot.n_dof
        dmin, dmax = _d(robot.dof_min), _d(robot.dof_max)
        keep += [dmin, dmax]
        d.dof_min, d.dof_max = _dp(dmin), _dp(dmax)
        ch = (ChainDesc * len(chains))(*chains); jt = (JointDesc * max(1, len(joints)))(*joints)
        lb = np.ascontiguousarray(link_body, dtype=np.int32); ll = np.ascontiguousarray(link_last, dtype=np.int32)
        keep += [ch, jt, lb, ll]
        d.nchains, d.chains = len(chains), ch
        d.njoints, d.joints = len(joints), jt
        d.nlinks, d.link_body, d.link_last_joint = len(link_body), _ip(lb), _ip(ll)
        d.parent_robot, d.parent_chain = parent_robot_id, max(robot.parent_chain, 0)
        off = parent_offset if parent_offset is not None else (np.array([1.0, 0, 0, 0]), np.zeros(3))
        d.parent_offset_q[:] = list(off[0]); d.parent_offset_t[:] = list(off[1])
        d.approach_q[:] = list(robot.approach[0]); d.approach_t[:] = list(robot.approach[1])
        if robot.eg is not None:
            eg, eo, en = _d(robot.eg), _d(robot.eg_origin), _d(robot.eg_norm)
            keep += [eg, eo, en]
            d.neg, d.eg, d.eg_origin, d.eg_norm = robot.eg.shape[0], _dp(eg), _dp(eo), _dp(en)
        else:
            d.neg = 0
        vcs = []
        for v in robot.vcontacts:
            vd = VContactDesc()
            vd.body = body_ids[v.body]
            vd.loc[:] = list(v.loc); vd.normal[:] = list(v.normal)
            vcs.append(vd)
        vc = (VContactDesc * max(1, len(vcs)))(*vcs)
        keep.append(vc)
        d.nvcontacts, d.vcontacts = len(vcs), vc
        rid = C.c_int(-1)
        self._ck(self.L.b2c_robot_define(self.h, C.byref(d), C.byref(rid)))
        return rid.value

    def robot_bodies(self, robot: int) -> List[int]:
        out = np.zeros(256, dtype=np.int32); n = C.c_int(0)
        self._ck(self.L.b2c_robot_bodies(self.h, robot, _ip(out), 256, C.byref(n)))
        return [int(v) for v in out[: n.value]]

    def robot_pairs(self, robot: int) -> List[Tuple[int, int]]:
        out = np.zeros(2 * 65536, dtype=np.int32); n = C.c_int(0)
        self._ck(self.L.b2c_robot_pairs(self.h, robot, _ip(out), 65536, C.byref(n)))
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n.value)]

    # --- best-row exchange between GPUs (no collective) -------------------------------------------------------------------
    def exchange_create(self, rank: int, world: int, rows: int, row_floats: int, depth: int = 4):
        """returns (64-byte IPC handle as bytes, own ring device pointer)"""
        h = (C.c_ubyte * 64)()
        ring = C.c_void_p()
        self._ck(self.L.b2c_exchange_create(self.h, rank, world, depth, rows, row_floats, h, C.byref(ring)))
        self._xshape = (world, rows, row_floats)
        return bytes(h), ring.value

    def exchange_connect(self, ipc_handles: bytes = None, peer_rings=None, peer_devices=None):
        if peer_rings is not None:
            arr = (C.c_void_p * len(peer_rings))(*peer_rings)
            dev = None if peer_devices is None else (C.c_int * len(peer_devices))(*peer_devices)
            self._ck(self.L.b2c_exchange_connect(self.h, None, arr, dev))
        else:
            buf = (C.c_ubyte * len(ipc_handles)).from_buffer_copy(ipc_handles)
            self._ck(self.L.b2c_exchange_connect(self.h, buf, None, None))

    def exchange_collect(self, rows_host_ptr: int, steps_host_ptr: int = 0, asynchronous=False):
        """newest complete rows of every rank -> host memory [world][rows][row_floats] (+ steps [world])"""
        self._ck(self.L.b2c_exchange_collect(self.h, rows_host_ptr, steps_host_ptr or None, int(bool(asynchronous))))

    def exchange_destroy(self):
        self._ck(self.L.b2c_exchange_destroy(self.h))

    def check_overflow(self):
        """synchronise and raise if a device-pointer evaluation overflowed an internal queue since the last check"""
        self._ck(self.L.b2c_check_overflow(self.h))

    def last_stats(self) -> dict:
        st = (C.c_ulonglong * 8)()
        self.L.b2c_last_stats(self.h, st)
        names = ["box_tests", "tri_tests", "point_box_tests", "point_tri_tests", "ambiguous", "dist_box_tests",
                 "dist_tri_tests", "recheck_hits"]
        return {k: int(st[i]) for i, k in enumerate(names)}

    def eval_batch(self, robot: int, obj: int, states, input_mode=B2C_INPUT_RAW, ref_tran=None, live=False,
                   pair_mask=False, link_dist=False, body_tf=False, position_params=None):
        """Host-buffer batched evaluation.  states: (npose, stride) float32.
        Returns dict(legal, energy[, pair_mask, link_dist, body_tf])."""
        st = np.ascontiguousarray(states, dtype=np.float64 if input_mode == B2C_INPUT_JOINTS else np.float32)
        npose, stride = st.shape
        a = EvalArgs()
        a.robot, a.object, a.npose, a.input_mode = robot, obj, npose, input_mode
        a.states, a.stride = st.ctypes.data, stride
        rq, rt = ref_tran if ref_tran is not None else (np.array([1.0, 0, 0, 0]), np.zeros(3))
        a.ref_q[:] = list(rq); a.ref_t[:] = list(rt)
        if position_params is not None:
            a.position_params[:] = [float(v) for v in position_params]
        flags = (B2C_EVAL_LIVE if live else 0) | (B2C_EVAL_PAIRMASK if pair_mask else 0) | (B2C_EVAL_LINKDIST if link_dist else 0)
        a.flags = flags
        legal = np.zeros(npose, dtype=np.uint8); energy = np.zeros(npose, dtype=np.float32)
        a.legal, a.energy = legal.ctypes.data, energy.ctypes.data
        res = {"legal": legal, "energy": energy}
        nb = len(self.robot_bodies(robot))
        if pair_mask:
            mw = max(1, (len(self.robot_pairs(robot)) + 63) // 64)
            pm = np.zeros((npose, mw), dtype=np.uint64)
            a.pair_mask = pm.ctypes.data
            res["pair_mask"] = pm
        if link_dist:
            nh = self._root_bodies(robot)
            ld = np.zeros((npose, nh), dtype=np.float32)
            a.link_dist = ld.ctypes.data
            res["link_dist"] = ld
        if body_tf:
            bt = np.zeros((npose, nb, 7), dtype=np.float64)
            a.body_tf = bt.ctypes.data
            res["body_tf"] = bt
        self._ck(self.L.b2c_eval_batch(self.h, C.byref(a)))
        return res

    def contacts_batch(self, robot: int, states, threshold=0.2, input_mode=B2C_INPUT_RAW, ref_tran=None, raw=False,
                       cap_contacts=None):
        """allContacts(threshold, interest = the robot's bodies) for every pose: returns dict(contacts (n, 13), pose (n,),
        pair (n,) indexing robot_pairs(robot), offsets (npose + 1,))."""
        st = np.ascontiguousarray(states, dtype=np.float64 if input_mode == B2C_INPUT_JOINTS else np.float32)
        npose, stride = st.shape
        cap = int(cap_contacts) if cap_contacts is not None else max(1024, 16 * npose)
        buf = np.empty((cap, 13), dtype=np.float64)              # b2c_contact = 13 doubles
        cpose = np.empty(cap, dtype=np.int32); cpair = np.empty(cap, dtype=np.int32); off = np.zeros(npose + 1, dtype=np.int32)
        n = C.c_int(0)
        a = ContactsBatchArgs()
        a.robot, a.npose, a.input_mode, a.states, a.stride = robot, npose, input_mode, st.ctypes.data, stride
        rq, rt = ref_tran if ref_tran is not None else (np.array([1.0, 0, 0, 0]), np.zeros(3))
        a.ref_q[:] = list(rq); a.ref_t[:] = list(rt)
        a.threshold, a.raw = float(threshold), int(bool(raw))
        a.contacts, a.contact_pose, a.contact_pair, a.cap_contacts = buf.ctypes.data, cpose.ctypes.data, cpair.ctypes.data, cap
        a.ncontacts, a.pose_offsets = C.addressof(n), off.ctypes.data
        self._ck(self.L.b2c_contacts_batch(self.h, C.byref(a)))
        if n.value > cap:
            return self.contacts_batch(robot, states, threshold, input_mode, ref_tran, raw, cap_contacts=n.value)
        k = n.value
        # (views of this call's own buffers: the untouched tail of the capacity estimate never becomes resident)
        return {"contacts": buf[:k], "pose": cpose[:k], "pair": cpair[:k], "offsets": off}

    def move_dof_to_contacts(self, robot: int, states, desired, steps, dof_breakaway, stop_at_contact=False,
                             contact_threshold=0.2, njoints=None):
        """Robot::moveDOFToContacts for a batch: states (npose, 7 + ndof) float32 RAW rows (collision-free start).
        Returns dict(dof (npose, ndof), joints, status (B2C_AG_* flags), iters, stopped, rounds)."""
        st = np.ascontiguousarray(states, dtype=np.float32)
        npose, stride = st.shape
        des = np.ascontiguousarray(desired, dtype=np.float64)
        ndof = des.size
        stp = None if steps is None else np.ascontiguousarray(steps, dtype=np.float64)
        brk = np.ascontiguousarray(dof_breakaway, dtype=np.uint8)
        assert brk.size == ndof and (stp is None or stp.size == ndof)
        nj = int(njoints) if njoints is not None else 32
        dof = np.zeros((npose, ndof)); joints = np.zeros((npose, nj)); status = np.zeros((npose, 2), dtype=np.int32)
        stopped = np.zeros((npose, nj), dtype=np.uint8)
        rounds = C.c_int(0)
        a = AutoGraspArgs()
        a.robot, a.npose, a.states, a.stride = robot, npose, st.ctypes.data, stride
        a.desired, a.steps, a.dof_breakaway = des.ctypes.data, (None if stp is None else stp.ctypes.data), brk.ctypes.data
        a.stop_at_contact, a.contact_threshold, a.max_rounds = int(bool(stop_at_contact)), float(contact_threshold), 0
        a.dof_out = dof.ctypes.data
        if njoints is not None:
            a.joint_out, a.stopped_out = joints.ctypes.data, stopped.ctypes.data
        a.status = status.ctypes.data
        a.rounds = C.addressof(rounds)
        self._ck(self.L.b2c_autograsp_batch(self.h, C.byref(a)))
        return {"dof": dof, "joints": joints if njoints is not None else None, "status": status[:, 0].copy(),
                "iters": status[:, 1].copy(), "stopped": stopped if njoints is not None else None, "rounds": rounds.value}

    def auto_grasp(self, robot: int, robot_desc, states, speed_factor=1.0, stop_at_contact=False):
        """Hand::autoGrasp (static branch, reference src/robot.cpp:2427-2462) for a batch of RAW states; robot_desc is the
        scene's RobotDesc (DOF limits, types and <defaultVelocity>)."""
        vel = np.asarray(robot_desc.dof_velocity, dtype=np.float64)
        desired = np.where(speed_factor * vel >= 0, robot_desc.dof_max, robot_desc.dof_min).astype(np.float64)
        steps = vel * speed_factor * AUTO_GRASP_TIME_STEP
        brk = [{"r": 0, "b": 1, "c": 2}[t] for t in robot_desc.dof_type]
        return self.move_dof_to_contacts(robot, states, desired, steps, brk, stop_at_contact, njoints=robot_desc.n_joints)

    def _root_bodies(self, robot: int) -> int:
        return getattr(self, "_nroot", {}).get(robot, len(self.robot_bodies(robot)))

    def eval_batch_device(self, robot: int, obj: int, states_ptr: int, npose: int, stride: int, legal_ptr: int,
                          energy_ptr: int, input_mode=B2C_INPUT_RAW, ref_tran=None, live=False):
        """Device-pointer variant: inputs/outputs already resident in HBM; asynchronous on ``stream``."""
        a = EvalArgs()
        a.robot, a.object, a.npose, a.input_mode = robot, obj, npose, input_mode
        a.states, a.stride = states_ptr, stride
        rq, rt = ref_tran if ref_tran is not None else (np.array([1.0, 0, 0, 0]), np.zeros(3))
        a.ref_q[:] = list(rq); a.ref_t[:] = list(rt)
        a.flags = B2C_EVAL_DEVICE_PTRS | (B2C_EVAL_LIVE if live else 0)
        a.legal, a.energy = legal_ptr, energy_ptr
        self._ck(self.L.b2c_eval_batch(self.h, C.byref(a)))


def _contact_array(rows):
    """(n, 13) rows -> (ctypes Contact array, n); one memcpy, b2c_contact is 13 consecutive doubles"""
    rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 13)
    n = len(rows)
    arr = (Contact * max(1, n))()
    if n:
        C.memmove(arr, rows.ctypes.data, 13 * 8 * n)
    return arr, n


def check_contact_normals(rows, tran1, tran2):
    """checkContactNormals over contact rows [n][13]; returns (rows with flipped normals, number changed).  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)
    q1, t1, q2, t2 = [np.ascontiguousarray(x, dtype=np.float64) for x in (tran1[0], tran1[1], tran2[0], tran2[1])]
    dp = C.POINTER(C.c_double)
    nf = C.c_int(0)
    rc = L.b2c_check_contact_normals(arr, n, q1.ctypes.data_as(dp), t1.ctypes.data_as(dp), q2.ctypes.data_as(dp), t2.ctypes.data_as(dp), C.byref(nf))
    if rc:
        raise B2CError(f"b2c_check_contact_normals failed ({rc})")
    return _contacts_to_array(arr, n), nf.value


def contact_frames(rows, side=1):
    """Contact::Contact frames of contact rows [n][13]: (n, 7) = q wxyz, t.  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)
    out = np.zeros((n, 7))
    rc = L.b2c_contact_frames(arr, n, side, out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise B2CError(f"b2c_contact_frames failed ({rc})")
    return out


def point_contact_wrenches(rows, side, cof, cog, max_radius):
    """PCWF friction-cone boundary wrenches of contact rows [n][13]: (n, 8, 6) = force, torque.  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)
    out = np.zeros((n, 8, 6))
    g = np.ascontiguousarray(cog, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    rc = L.b2c_point_contact_wrenches(arr, n, side, float(cof), g.ctypes.data_as(dp), float(max_radius), out.ctypes.data_as(dp))
    if rc:
        raise B2CError(f"b2c_point_contact_wrenches failed ({rc})")
    return out


def _rows_of(arr, n):
    return _contacts_to_array(arr, n)


def _ragged(lists):
    """list of (k_i, 3) point arrays -> (flat xyz float64, offsets int32[n + 1])"""
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(np.asarray(l).reshape(-1, 3))
    flat = np.zeros((max(int(off[-1]), 1), 3))
    for i, l in enumerate(lists):
        flat[off[i]:off[i + 1]] = np.asarray(l, dtype=np.float64).reshape(-1, 3)
    return np.ascontiguousarray(flat), off


def soft_neighborhood_radius(youngs1, youngs2):
    """findSoftNeighborhoods' bodyRegion radius in mm (contactSetting.cpp:129-133).  Host only."""
    return float(load_library().b2c_soft_neighborhood_radius(float(youngs1), float(youngs2)))


def soft_contact_merge(rows, tran1, tran2, nghbd1, nghbd2):
    """mergeSoftNeighborhoods: returns (rows', nghbd1', nghbd2', merges); nghbd* are lists of (k, 3) arrays.  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)
    q1, t1, q2, t2 = [np.ascontiguousarray(x, dtype=np.float64) for x in (tran1[0], tran1[1], tran2[0], tran2[1])]
    x1, o1 = _ragged(nghbd1); x2, o2 = _ragged(nghbd2)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    nn, nm = C.c_int(n), C.c_int(0)
    rc = L.b2c_soft_contact_merge(arr, C.byref(nn), q1.ctypes.data_as(dp), t1.ctypes.data_as(dp), q2.ctypes.data_as(dp),
                                  t2.ctypes.data_as(dp), x1.ctypes.data_as(dp), o1.ctypes.data_as(ip), x2.ctypes.data_as(dp),
                                  o2.ctypes.data_as(ip), C.byref(nm))
    if rc:
        raise B2CError(f"b2c_soft_contact_merge failed ({rc})")
    k = nn.value
    return (_rows_of(arr, k), [x1[o1[i]:o1[i + 1]].copy() for i in range(k)], [x2[o2[i]:o2[i + 1]].copy() for i in range(k)],
            nm.value)


def soft_contact_fit(rows, side, nghbd):
    """SoftContact surface fit per contact: (n, 6) = a, b, c, r1, r2, rotAngle.  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)
    x, o = _ragged(nghbd)
    assert len(nghbd) == n
    fit = (SoftFit * max(n, 1))()
    rc = L.b2c_soft_contact_fit(arr, n, side, x.ctypes.data_as(C.POINTER(C.c_double)), o.ctypes.data_as(C.POINTER(C.c_int)), fit)
    if rc:
        raise B2CError(f"b2c_soft_contact_fit failed ({rc})")
    return np.array([[f.a, f.b, f.c, f.r1, f.r2, f.rot_angle] for f in fit[:n]]).reshape(-1, 6)


def soft_contact_wrenches(rows, side, fit_this, fit_mate, q_this, q_mate, youngs, cof, cog, max_radius):
    """SFCL friction ellipsoid -> boundary wrenches: ((n, 20, 6) force/torque, (n, 6) major, minor, relPhi, r1', r2',
    torque/N).  Host only."""
    L = load_library()
    arr, n = _contact_array(rows)

    def fits(a):
        a = np.asarray(a, dtype=np.float64).reshape(-1, 6)
        f = (SoftFit * max(n, 1))()
        for i in range(n):
            f[i].a, f[i].b, f[i].c, f[i].r1, f[i].r2, f[i].rot_angle = [float(v) for v in a[i]]
        return f
    out = np.zeros((n, 20, 6)); ell = np.zeros((n, 6))
    qa, qb, g = [np.ascontiguousarray(v, dtype=np.float64) for v in (q_this, q_mate, cog)]
    dp = C.POINTER(C.c_double)
    rc = L.b2c_soft_contact_wrenches(arr, n, side, fits(fit_this), fits(fit_mate), qa.ctypes.data_as(dp), qb.ctypes.data_as(dp),
                                     float(youngs), float(cof), g.ctypes.data_as(dp), float(max_radius),
                                     out.ctypes.data_as(dp), ell.ctypes.data_as(dp))
    if rc:
        raise B2CError(f"b2c_soft_contact_wrenches failed ({rc})")
    return out, ell


class Annealer:
    """K annealing chains on the device (``b2c_anneal_*``): host mirror of ``SimAnnPlanner``'s loop
    (reference src/EGPlanner/simAnnPlanner.cpp:111-148) for a batch of independent chains."""

    def __init__(self, engine: "Engine", robot: int, obj: int, variables, params, initial_states, ref_tran,
                 seed=1234, chain_offset=0, live=False):
        """variables: planner.SearchVariables; params: planner.SimAnnParams"""
        self.engine = engine
        st = np.ascontiguousarray(initial_states, dtype=np.float32)
        self.nchain, self.nvar = st.shape
        d = AnnealDesc()
        d.robot, d.object, d.nchain, d.nvar = robot, obj, self.nchain, self.nvar
        self._keep = [np.ascontiguousarray(variables.vmin, dtype=np.float64), np.ascontiguousarray(variables.vmax, dtype=np.float64),
                      np.ascontiguousarray(variables.vjump, dtype=np.float64), np.ascontiguousarray(variables.circular, dtype=np.uint8)]
        dp = C.POINTER(C.c_double)
        d.vmin, d.vmax, d.vjump = (self._keep[i].ctypes.data_as(dp) for i in range(3))
        d.circular = self._keep[3].ctypes.data_as(C.POINTER(C.c_ubyte))
        d.ref_q[:] = list(ref_tran[0]); d.ref_t[:] = list(ref_tran[1])
        d.YC, d.HC, d.NBR_ADJ, d.ERR_ADJ, d.T0 = params.YC, params.HC, params.NBR_ADJ, params.ERR_ADJ, params.T0
        d.YDIMS, d.HDIMS, d.K0 = params.YDIMS, params.HDIMS, params.K0
        d.seed, d.chain_offset, d.flags = seed, chain_offset, (B2C_EVAL_LIVE if live else 0)
        aid = C.c_int(-1)
        engine._ck(engine.L.b2c_anneal_create(engine.h, C.byref(d), st.ctypes.data_as(C.POINTER(C.c_float)), C.byref(aid)))
        self.id = aid.value

    def step(self, nsteps=1):
        self.engine._ck(self.engine.L.b2c_anneal_step(self.engine.h, self.id, nsteps))

    def read(self):
        n, nv = self.nchain, self.nvar
        out = dict(cur=np.zeros((n, nv), np.float32), cur_energy=np.zeros(n, np.float32), best=np.zeros((n, nv), np.float32),
                   best_energy=np.zeros(n, np.float32), k=np.zeros(n, np.int32))
        tot = (C.c_longlong * 4)()
        self.engine._ck(self.engine.L.b2c_anneal_read(self.engine.h, self.id, out["cur"].ctypes.data, out["cur_energy"].ctypes.data,
                                                      out["best"].ctypes.data, out["best_energy"].ctypes.data, out["k"].ctypes.data, tot))
        out.update(jumps=int(tot[0]), legal_neighbours=int(tot[1]), steps=int(tot[2]), evaluated=int(tot[3]))
        return out

    def device_ptrs(self):
        p = [C.c_void_p() for _ in range(4)]
        self.engine._ck(self.engine.L.b2c_anneal_device_ptrs(self.engine.h, self.id, *[C.byref(x) for x in p]))
        return dict(best=p[0].value, best_energy=p[1].value, cur=p[2].value, cur_energy=p[3].value)

    def best_rows(self, k: int, rows_device_ptr: int):
        """this context's k best (state || energy) rows into device memory [k][nvar + 1]; asynchronous"""
        self.engine._ck(self.engine.L.b2c_anneal_best_rows(self.engine.h, self.id, k, rows_device_ptr))

    def publish(self, rows_device_ptr: int = 0):
        """fused best-list update + store of this rank's best rows into every peer's ring (b2c_anneal_publish); asynchronous"""
        self.engine._ck(self.engine.L.b2c_anneal_publish(self.engine.h, self.id, rows_device_ptr or None))

    def torch_views(self, device):
        """torch tensors aliasing the annealer's device rows (no copy): best, best_energy, cur, cur_energy"""
        import torch

        class _Dev:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}
        p = self.device_ptrs()
        shp = {"best": (self.nchain, self.nvar), "best_energy": (self.nchain,), "cur": (self.nchain, self.nvar), "cur_energy": (self.nchain,)}
        return {k: torch.as_tensor(_Dev(p[k], shp[k]), device=device) for k in shp}

    def propose(self):
        out = np.zeros((self.nchain, self.nvar), np.float32)
        self.engine._ck(self.engine.L.b2c_anneal_propose(self.engine.h, self.id, out.ctypes.data))
        return out

    def write(self, cur=None, cur_energy=None, k=None, step=-1):
        a = [None if x is None else np.ascontiguousarray(x, dtype=t) for x, t in ((cur, np.float32), (cur_energy, np.float32), (k, np.int32))]
        self.engine._ck(self.engine.L.b2c_anneal_write(self.engine.h, self.id, *[None if x is None else x.ctypes.data for x in a], step))

    def close(self):
        if self.id >= 0 and self.engine.h:
            self.engine.L.b2c_anneal_destroy(self.engine.h, self.id)
        self.id = -1


class SceneBinding:
    """A ``formats.xmlmodel.Scene`` loaded into an :class:`Engine`: bodies, transforms, the file-defined
    disabled pairs, the pairs found touching at load, and the robots."""

    def __init__(self, engine: Engine, scene, apply_touching=True):
        self.engine = engine
        self.scene = scene
        self.body_ids = {}
        for i, b in enumerate(scene.bodies):
            self.body_ids[i] = engine.add_body(b.tris)
            q, t = scene.body_tran[i]
            engine.set_body_transform(self.body_ids[i], q, t)
        for (a, b) in scene.disabled_pairs:
            engine.activate_pair(self.body_ids[a], self.body_ids[b], False)
        if apply_touching:
            for (a, b) in scene.touching_pairs:
                engine.activate_pair(self.body_ids[a], self.body_ids[b], False)
        self.robot_ids = {}
        engine._nroot = getattr(engine, "_nroot", {})
        for ri, r in enumerate(scene.robots):
            off = None
            parent_id = -1
            if r.parent >= 0:
                parent_id = self.robot_ids[r.parent]
                for (child, o) in scene.robots[r.parent].chains[r.parent_chain].children:
                    if child == ri:
                        off = o
            rid = engine.define_robot(r, self.body_ids, parent_id, off)
            self.robot_ids[ri] = rid
            engine._nroot[rid] = sum(len(c.link_bodies) for c in r.chains) + 1      # finger links + palm (no mount)
