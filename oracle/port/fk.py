"""TEST INFRASTRUCTURE ONLY -- oracle restatement (numpy, fp64) of GraspIt!'s forward kinematics
and of the search-state -> hand-pose mapping.  Never imported by the product path.

Reference:
  * KinematicChain::fwdKinematics        src/kinematicChain.cpp:543-560
      total = owner->getTran() % tran; per joint: total = total % joint->getTran(val);
      link l takes the running product after joint lastJoint[l]
  * DHTransform::getTran(theta,d)        include/graspit/joint.h:126-130
      AXIS_ANGLE_ROTATION(theta, z) % TRANSLATION(0,0,d) % (TRANSLATION(a,0,0) % AXIS_ANGLE_ROTATION(alpha, x))
      with tr4TimesTr3 = tr3 % tr4       src/joint.cpp:68-88
  * Revolute/PrismaticJoint::getTran     include/graspit/joint.h:396-398 (theta = val + c / d = val + c)
  * Rigid/BreakAwayDOF::accumulateMove   src/dof.cpp:500-514,577-620: joint = dof * couplingRatio
      (stateless evaluation: the 1e-5 "already there" hysteresis and break-away memory do not
      apply because forceDOFVals resets the DOF first, include/graspit/robot.h:654-664)
  * Robot::simpleSetTran / attachRobot   include/graspit/robot.h:601-605, src/kinematicChain.cpp:524-526:
      child base = last link tran % childOffsetTran; mount piece moves with the child base
  * HandObjectState::execute             src/EGPlanner/searchState.cpp:515-527
  * PositionStateAA::getCoreTran         src/EGPlanner/searchStateImpl.cpp:156-168
  * PostureStateEigen::getHandDOF        src/EGPlanner/searchStateImpl.cpp:83-95
  * EigenGraspInterface::toDOFSpace      src/eigenGrasp.cpp:638-646 (mPInv = E^T, :567)
  * Robot::checkSetDOFVals               include/graspit/robot.h:698-709

The robot argument is duck-typed (plain data: chains[].tran/.joints/.link_bodies/.last_joint/.children,
base_body, mount_body, approach, eg, eg_origin, eg_norm, dof_min, dof_max).
"""
import numpy as np

from . import transf_np as T

_Z = np.array([0.0, 0.0, 1.0])
_X = np.array([1.0, 0.0, 0.0])


def _b(tr, n):
    """broadcast a single (q,t) to batch n"""
    q = np.broadcast_to(np.asarray(tr[0], dtype=np.float64).reshape(-1, 4), (n, 4)).copy()
    t = np.broadcast_to(np.asarray(tr[1], dtype=np.float64).reshape(-1, 3), (n, 3)).copy()
    return q, t


def joint_values(robot, dofs):
    """(N,nDOF) -> list per chain of (N,nJoints) joint values: joint = ratio * dof."""
    out = []
    for c in robot.chains:
        jv = np.stack([dofs[:, j.dof] * j.ratio for j in c.joints], axis=1)
        out.append(jv)
    return out


def joint_transf(j, val):
    n = val.shape[0]
    if j.revolute:
        theta = val + j.theta          # j.theta holds c (DH theta at value 0)
        d = np.full(n, j.d)
    else:
        theta = np.full(n, j.theta)
        d = val + j.d
    tr1 = T.axis_angle_rotation(theta, _Z)
    tr2 = T.translation(np.stack([np.zeros(n), np.zeros(n), d], axis=1))
    tr3 = T.translation(np.array([[j.a, 0.0, 0.0]]))
    tr4 = T.axis_angle_rotation(np.array([j.alpha]), _X)
    tr4tr3 = T.compose(tr3, tr4)
    return T.compose(T.compose(tr1, tr2), _b(tr4tr3, n))


def forward_kinematics(robots, ridx, base, dofs_by_robot, out=None):
    """Poses of every body of robot ``ridx`` (and of robots attached to it).

    base: (q (N,4), t (N,3)) world-from-base; dofs_by_robot: {robot index: (N,nDOF)}.
    Returns {scene body index: (q,t)}."""
    if out is None:
        out = {}
    robot = robots[ridx]
    n = base[0].shape[0]
    out[robot.base_body] = base
    if getattr(robot, "mount_body", -1) >= 0:
        out[robot.mount_body] = base
    dofs = dofs_by_robot[ridx]
    jvs = joint_values(robot, dofs)
    for c, jv in zip(robot.chains, jvs):
        total = T.compose(base, _b(c.tran, n))
        l = 0
        nl = len(c.link_bodies)
        for j, jd in enumerate(c.joints):
            total = T.compose(total, joint_transf(jd, jv[:, j]))
            while l < nl and c.last_joint[l] == j:
                out[c.link_bodies[l]] = total
                l += 1
                break
        for (child, off) in getattr(c, "children", []):
            last = out[c.link_bodies[-1]]
            forward_kinematics(robots, child, T.compose(last, _b(off, n)), dofs_by_robot, out)
    return out


def eigen_to_dof(robot, amps):
    """PostureStateEigen::getHandDOF with a rigid interface: dof = (E^T amp) * norm + origin, clamped."""
    amps = np.asarray(amps, dtype=np.float64)
    x = amps @ robot.eg                       # (N,nEG) @ (nEG,nDOF)
    dof = x * robot.eg_norm[None, :] + robot.eg_origin[None, :]
    return np.clip(dof, robot.dof_min[None, :], robot.dof_max[None, :])


def clamp_dofs(robot, dofs):
    return np.clip(dofs, robot.dof_min[None, :], robot.dof_max[None, :])


def aa_state_to_base(robot, ref_tran, pos6):
    """HandObjectState::execute for SPACE_AXIS_ANGLE: base = mRefTran % coreTran,
    coreTran = TRANSLATION(tx,ty,tz) % AXIS_ANGLE_ROTATION(alpha, (sin th cos ph, sin th sin ph, cos th))
               % approachTran^-1."""
    pos6 = np.asarray(pos6, dtype=np.float64)
    n = pos6.shape[0]
    tx, ty, tz, th, ph, al = [pos6[:, i] for i in range(6)]
    axis = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    core = T.compose(T.translation(np.stack([tx, ty, tz], axis=1)), T.axis_angle_rotation(al, axis))
    core = T.compose(core, _b(T.inverse(_b(robot.approach, 1)), n))
    return T.compose(_b(ref_tran, n), core)


def complete_state_to_base(ref_tran, pos7):
    """SPACE_COMPLETE (searchStateImpl.cpp:125-135): Tx Ty Tz Qw Qx Qy Qz."""
    pos7 = np.asarray(pos7, dtype=np.float64)
    n = pos7.shape[0]
    core = T.make(pos7[:, 3:7], pos7[:, 0:3])
    return T.compose(_b(ref_tran, n), core)


def _quat_from_matrix(m):
    """rotation matrix (N,3,3) -> unit quaternion wxyz, Eigen's trace-branch algorithm (what transf::set(mat3, vec3) stores,
    include/graspit/transform.h:52-57)"""
    m = np.asarray(m, dtype=np.float64)
    out = np.zeros((m.shape[0], 4))
    for k in range(m.shape[0]):
        M = m[k]
        t = np.trace(M)
        c = np.zeros(4)                      # x y z w
        if t > 0:
            t = np.sqrt(t + 1.0)
            c[3] = 0.5 * t
            t = 0.5 / t
            c[0] = (M[2, 1] - M[1, 2]) * t; c[1] = (M[0, 2] - M[2, 0]) * t; c[2] = (M[1, 0] - M[0, 1]) * t
        else:
            i = 0
            if M[1, 1] > M[0, 0]:
                i = 1
            if M[2, 2] > M[i, i]:
                i = 2
            j = (i + 1) % 3; kk = (j + 1) % 3
            t = np.sqrt(M[i, i] - M[j, j] - M[kk, kk] + 1.0)
            c[i] = 0.5 * t
            t = 0.5 / t
            c[3] = (M[kk, j] - M[j, kk]) * t; c[j] = (M[j, i] + M[i, j]) * t; c[kk] = (M[kk, i] + M[i, kk]) * t
        out[k] = [c[3], c[0], c[1], c[2]]
    return out / np.linalg.norm(out, axis=1, keepdims=True)


def _unit(v):
    n = np.linalg.norm(v, axis=1, keepdims=True)
    return np.where(n > 0, v / np.where(n > 0, n, 1.0), v)       # Eigen normalized(): unchanged at zero


def ellipsoid_state_to_base(robot, ref_tran, pos4, abc=(80.0, 80.0, 160.0)):
    """SPACE_ELLIPSOID (PositionStateEllipsoid::getCoreTran, searchStateImpl.cpp:218-252): beta gamma tau dist; the hand sits
    on the ellipsoid (a, b, c), z along the (inward) normal, rolled by tau, `dist` back along the un-normalised normal."""
    pos4 = np.asarray(pos4, dtype=np.float64)
    n = pos4.shape[0]
    a, b, c = abc
    be, ga, tau, dist = [pos4[:, i] for i in range(4)]
    p = np.stack([a * np.cos(be) * np.cos(ga), b * np.cos(be) * np.sin(ga), c * np.sin(be)], 1)
    n1 = np.stack([-a * np.sin(be) * np.cos(ga), -b * np.sin(be) * np.sin(ga), c * np.cos(be)], 1)
    n2 = np.stack([-a * np.cos(be) * np.sin(ga), b * np.cos(be) * np.cos(ga), np.zeros(n)], 1)
    normal = np.cross(_unit(n1), _unit(n2))
    xdir = np.tile(np.array([[1.0, 0.0, 0.0]]), (n, 1))
    ydir = np.cross(normal, _unit(xdir))
    xdir = np.cross(ydir, normal)
    r = np.stack([_unit(xdir), _unit(ydir), _unit(normal)], 2)          # columns
    hand = (_quat_from_matrix(r), p - dist[:, None] * normal)
    zrot = (np.stack([np.cos(0.5 * tau), np.zeros(n), np.zeros(n), np.sin(0.5 * tau)], 1), np.zeros((n, 3)))   # plain AngleAxisd: no snap
    hand = T.compose(hand, T.make(zrot[0], zrot[1]))
    core = T.compose(hand, _b(T.inverse(_b(robot.approach, 1)), n))
    return T.compose(_b(ref_tran, n), core)


def approach_state_to_base(robot, ref_tran, pos3):
    """SPACE_APPROACH (PositionStateApproach::getCoreTran, searchStateImpl.cpp:266-274): dist, wrist 1, wrist 2:
    approach % [ R(ry, y) % R(rx, x) % T(0, 0, dist) ] % approach^-1."""
    pos3 = np.asarray(pos3, dtype=np.float64)
    n = pos3.shape[0]
    dist, rx, ry = pos3[:, 0], pos3[:, 1], pos3[:, 2]
    hand = T.translation(np.stack([np.zeros(n), np.zeros(n), dist], 1))
    hand = T.compose(T.compose(T.axis_angle_rotation(ry, np.array([0.0, 1.0, 0.0])), T.axis_angle_rotation(rx, np.array([1.0, 0.0, 0.0]))), hand)
    ap = _b(robot.approach, n)
    core = T.compose(T.compose(ap, hand), _b(T.inverse(_b(robot.approach, 1)), n))
    return T.compose(_b(ref_tran, n), core)
