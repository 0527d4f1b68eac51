"""TEST INFRASTRUCTURE ONLY -- CPU restatement of GraspIt!'s static DOF closing (autograsp), one pose at a time,
on top of the reference's own collision backend (oracle/_ref: allCollisions / bodyToBodyDistance are the verbatim
compiled GraspitCollision).  Never imported by the product path.

robot.cpp / dof.cpp are Qt-bound and cannot be compiled here, so the control flow is restated:
  * Hand::autoGrasp                      src/robot.cpp:2427-2462  (static branch: desired = DOF max / min by the sign of
                                         speedFactor * defaultVelocity, step = defaultVelocity * speedFactor * 0.01)
  * Robot::moveDOFToContacts             src/robot.cpp:1708-1815  (step loop, 500-iteration failsafe)
  * Robot::jumpDOFToContact              src/robot.cpp:1593-1706
  * Robot::getJointValuesFromDOF         src/robot.cpp:1507-1560
  * Robot::interpolateJoints             src/robot.cpp:1418-1472  (bisection on getDist of the colliding pairs)
  * Robot::stopJointsFromLink            src/robot.cpp:1474-1489
  * World::findContacts(CollisionReport&) src/world.cpp:1682-1707 (pairs farther than Contact::THRESHOLD are dropped)
  * RigidDOF / BreakAwayDOF              src/dof.cpp:497-514, 540-648 (accumulateMove, updateVal, getJointValues)
Not restated: Link::contactsPreventMotion inside getJointValuesFromDOF.  Starting from a hand without contacts (the
state HandObjectState::execute leaves: Body::setTran breaks the contacts of every link it moves, body.cpp:928-931) it
cannot fire: contacts are only added by findContacts to links whose upstream moving joints are stopped in the same
call, so the infinitesimal motion of every link that holds a contact is the identity (kinematicChain.cpp:610-619) and
Contact::preventsMotion (contact.cpp:390-394) compares 0 < -1e-9.
The reference holds no golden values for autograsp: **parity unpinned** beyond the collision backend itself.
"""
from __future__ import annotations

import numpy as np

from . import fk as ofk
from . import transf_np as T

CONTACT_THRESHOLD = 0.2          # Contact::THRESHOLD, src/contact/contact.cpp:58
AUTO_GRASP_TIME_STEP = 0.01      # Robot::AUTO_GRASP_TIME_STEP, src/robot.cpp:68


class _Dof:
    def __init__(self, kind, q, joints, ratios):
        self.kind, self.q, self.joints, self.ratios = kind, float(q), joints, ratios
        self.in_break = [False] * len(joints)
        self.break_val = [-10.0] * len(joints)

    def joint_values(self, jv):
        for i, (j, r) in enumerate(zip(self.joints, self.ratios)):
            if self.kind == "b" and self.in_break[i] and self.q > self.break_val[i]:
                jv[j] = self.break_val[i] * r
            else:
                jv[j] = self.q * r

    def accumulate_move(self, q1, jv, cur_jv, stopped):
        if abs(self.q - q1) < 1.0e-5:
            return False
        if self.kind == "c":                                   # CompliantDOF::accumulateMove (dof.cpp:830-856)
            movement = False
            for j, r in zip(self.joints, self.ratios):
                new, old = q1 * r, cur_jv[j]
                if new > old and (stopped[j] & 1):
                    continue
                if new < old and (stopped[j] & 2):
                    continue
                jv[j] = new
                movement = True
            return movement
        if self.kind != "b":                                   # RigidDOF
            if any(stopped[j] for j in self.joints):
                return False
            for j, r in zip(self.joints, self.ratios):
                jv[j] = q1 * r
            return True
        for j in self.joints:                                  # BreakAwayDOF
            if stopped[j] and q1 < self.q:
                return False
        movement = False
        for i, (j, r) in enumerate(zip(self.joints, self.ratios)):
            if self.in_break[i] and q1 > self.break_val[i]:
                continue
            new, old = q1 * r, cur_jv[j]
            if new > old and (stopped[j] & 1):
                continue
            if new < old and (stopped[j] & 2):
                continue
            jv[j] = new
            movement = True
        return movement

    def update_val(self, q1, cur_jv):
        if self.kind == "b":
            for i, (j, r) in enumerate(zip(self.joints, self.ratios)):
                jval = cur_jv[j] / r
                if self.in_break[i]:
                    if q1 < self.break_val[i] - 1.0e-5:
                        self.in_break[i] = False
                elif jval < q1 - 1.0e-5:
                    self.in_break[i] = True
                    self.break_val[i] = jval
        self.q = q1


class AutoGraspOracle:
    """one robot (no attached children) inside a reference world"""

    def __init__(self, scene, ridx, world, ids):
        self.sc, self.r, self.w, self.ids = scene, scene.robots[ridx], world, ids
        assert not any(c.children for c in self.r.chains), "autograsp oracle: single-robot trees only"
        self.inv = {v: k for k, v in ids.items()}
        self.jnum, self.jdesc = [], []
        for c in self.r.chains:
            self.jnum.append([len(self.jdesc) + k for k in range(len(c.joints))])
            self.jdesc += list(c.joints)
        self.nj = len(self.jdesc)
        self.ncoll = self.ndist = 0

    # -- kinematics -------------------------------------------------------------------------------------------
    def set_joints(self, base, jv):
        """KinematicChain::fwdKinematics from explicit joint values -> link poses into the reference world"""
        r = self.r
        # transf(Quaternion, vec3) normalises its rotation (transform.h:46-48): a base quaternion that went through fp32 is
        # 4e-8 off unit length, and an unnormalised one would leak into every link translation below
        b = T.make(np.asarray(base[0], dtype=np.float64)[None], np.asarray(base[1], dtype=np.float64)[None])
        self.last_tf = {r.base_body: (b[0][0].copy(), b[1][0].copy())}      # scene body -> (q wxyz, t) as pushed to the world
        self.w.set_transform(self.ids[r.base_body], b[0][0], b[1][0])
        if r.mount_body >= 0:
            self.last_tf[r.mount_body] = self.last_tf[r.base_body]
            self.w.set_transform(self.ids[r.mount_body], b[0][0], b[1][0])
        for ci, c in enumerate(r.chains):
            total = T.compose(b, ofk._b(c.tran, 1))
            l = 0
            for j, jd in enumerate(c.joints):
                total = T.compose(total, ofk.joint_transf(jd, np.array([jv[self.jnum[ci][j]]])))
                if l < len(c.link_bodies) and c.last_joint[l] == j:
                    self.w.set_transform(self.ids[c.link_bodies[l]], total[0][0], total[1][0])
                    self.last_tf[c.link_bodies[l]] = (total[0][0].copy(), total[1][0].copy())
                    l += 1
        self.jv = list(jv)

    def _new_dofs(self, dof0):
        out = []
        for d in range(self.r.n_dof):
            js = [k for k, jd in enumerate(self.jdesc) if jd.dof == d]
            out.append(_Dof(self.r.dof_type[d], dof0[d], js, [self.jdesc[k].ratio for k in js]))
        return out

    # -- Robot::interpolateJoints -----------------------------------------------------------------------------
    def _interpolate(self, base, initial, final, col):
        t, deltat, min_dist = 0.0, 1.0, 1.0e10
        while deltat > 1.0e-20 and t >= 0:
            deltat *= 0.5
            self.set_joints(base, [t * f + (1.0 - t) * i for i, f in zip(initial, final)])
            min_dist = 1.0e10
            for (a, b) in col:
                self.ndist += 1
                d, _, _ = self.w.body_distance(self.ids[a], self.ids[b])
                min_dist = min(min_dist, d)
                if min_dist <= 0:
                    t -= deltat
                    break
            if min_dist > 0:
                if min_dist < 0.5 * CONTACT_THRESHOLD:
                    break
                t += deltat
        ok = not (deltat < 1.0e-20 or t < 0)
        return ok, t

    def _stop_joints_from_link(self, body, desired_jv, stopped):
        bd = self.sc.bodies[body]
        if bd.chain < 0 or bd.link < 0:
            return
        c = self.r.chains[bd.chain]
        for j in range(c.last_joint[bd.link], -1, -1):
            k = self.jnum[bd.chain][j]
            if desired_jv[k] > self.jv[k]:
                stopped[k] |= 1
            elif desired_jv[k] < self.jv[k]:
                stopped[k] |= 2

    # -- Robot::jumpDOFToContact ------------------------------------------------------------------------------
    def _jump(self, base, dofs, desired, stopped):
        initial_dof = [d.q for d in dofs]
        initial_jv = list(self.jv)
        new_jv = list(self.jv)
        new_dof = [0.0] * len(dofs)
        moving = False
        for d, dof in enumerate(dofs):                        # getJointValuesFromDOF (contactsPreventMotion: see header)
            if dof.accumulate_move(desired[d], new_jv, self.jv, stopped):
                moving = True
                new_dof[d] = desired[d]
            else:
                new_dof[d] = dof.q
        if not moving:
            return False, 0
        interest = []
        for ci, c in enumerate(self.r.chains):
            for l, body in enumerate(c.link_bodies):
                if not stopped[self.jnum[ci][c.last_joint[l]]]:
                    interest.append(self.ids[body])
        self.set_joints(base, new_jv)
        late = []
        while True:
            self.ncoll += 1
            _, pairs = self.w.all_collisions(fast=False, interest=interest)
            col = [(self.inv[a], self.inv[b]) for a, b in pairs]
            if not col:
                break
            new_jv = list(self.jv)
            ok, t = self._interpolate(base, initial_jv, new_jv, col)
            if not ok:
                return None, 0                                   # interpolation error: DOF values are not updated
            late = list(col)
            new_dof = [n * t + i * (1.0 - t) for n, i in zip(new_dof, initial_dof)]
        kept = []
        for pr in late:                                          # World::findContacts: duplicates, then the distance filter
            if pr in kept:
                continue
            self.ndist += 1
            if self.w.body_distance(self.ids[pr[0]], self.ids[pr[1]])[0] > CONTACT_THRESHOLD:
                continue
            kept.append(pr)
        for (a, b) in kept:
            for body in (a, b):
                if self.sc.bodies[body].robot == self.sc.robots.index(self.r):
                    self._stop_joints_from_link(body, new_jv, stopped)
        for d, dof in enumerate(dofs):                           # Robot::updateDofVals
            dof.update_val(new_dof[d], self.jv)
        return True, len(kept)

    # -- Robot::moveDOFToContacts -----------------------------------------------------------------------------
    def move_dof_to_contacts(self, base, dof0, desired, steps, stop_at_contact=False):
        """returns dict(dof, joints, moved, iters, stopped, error)"""
        dofs = self._new_dofs(dof0)
        jv = [0.0] * self.nj
        for d in dofs:
            d.joint_values(jv)
        self.set_joints(base, jv)
        desired = [float(x) for x in desired]
        step = [0.0] * len(dofs)
        for i, d in enumerate(dofs):
            if steps is None:
                step[i] = desired[i] - d.q                       # WorldElement::ONE_STEP
            elif steps[i] != 0.0:
                step[i] = (-1.0 if desired[i] < d.q else 1.0) * abs(steps[i])
            else:
                step[i] = 0.0
                desired[i] = d.q
        stopped = [0] * self.nj
        moved, iters, error = False, 0, False
        while True:
            iters += 1
            new = []
            for i, d in enumerate(dofs):
                v = d.q + step[i]
                if step[i] > 0 and v > desired[i]:
                    v = desired[i]
                elif step[i] < 0 and v < desired[i]:
                    v = desired[i]
                new.append(v)
            ok, ncols = self._jump(base, dofs, new, stopped)
            if ok is None:
                error = True
                break
            if not ok:
                break
            moved = True
            if stop_at_contact and ncols != 0:
                break
            if iters > 500:
                break
        return dict(dof=np.array([d.q for d in dofs]), joints=np.array(self.jv), moved=moved, iters=iters,
                    stopped=np.array(stopped), error=error)

    def auto_grasp(self, base, dof0, speed_factor=1.0, stop_at_contact=False):
        desired, steps = auto_grasp_targets(self.r, speed_factor)
        return self.move_dof_to_contacts(base, dof0, desired, steps, stop_at_contact)


def compliant_static_ratios(robot, stiffness=None):
    """CompliantDOF::getStaticRatio (dof.cpp:779-795) for every joint of `robot` (RobotDesc), in the robot's joint order
    (chain-major = Joint::getNum): joints 0, 2, 4, 6 keep their coupling ratio; any other joint gets
    (ref / k) * tau1 * tau2 / (tau1 + tau2) with tau1 the coupling ratio and ref the spring stiffness of the DOF's first
    joint, tau2 / k its own (k == 0 -> ref).  Rigid and break-away DOFs: the coupling ratio (dof.cpp:484-487)."""
    joints = [jd for c in robot.chains for jd in c.joints]
    k_all = [1.0] * len(joints) if stiffness is None else list(stiffness)
    first = {}
    for jn, jd in enumerate(joints):
        first.setdefault(jd.dof, jn)
    out = []
    for jn, jd in enumerate(joints):
        if robot.dof_type[jd.dof] != "c" or jn in (0, 2, 4, 6):
            out.append(jd.ratio)
            continue
        ref = k_all[first[jd.dof]]
        k = k_all[jn] if k_all[jn] != 0.0 else ref
        tau2, tau1 = jd.ratio, joints[first[jd.dof]].ratio
        out.append((ref / k) * (tau1 * tau2) / (tau1 + tau2))
    return out


def with_static_ratios(scene, ridx=0, stiffness=None):
    """copy of `scene` whose JointDesc.ratio hold DOF::getStaticRatio (what include/b2c.h asks a caller to put into
    b2c_joint_desc.ratio for CompliantDOFs) and whose DOF limits follow DOF::updateMinMax (dof.cpp:92-119) on them"""
    import copy
    sc = copy.deepcopy(scene)
    r = sc.robots[ridx]
    ratios = compliant_static_ratios(r, stiffness)
    joints = [jd for c in r.chains for jd in c.joints]
    for jd, x in zip(joints, ratios):
        jd.ratio = x
    mn, mx = np.array(r.dof_min, dtype=np.float64), np.array(r.dof_max, dtype=np.float64)
    for d in range(r.n_dof):
        js = [jd for jd in joints if jd.dof == d]
        lo, hi = [], []
        for jd in js:
            a, b = jd.min / jd.ratio, jd.max / jd.ratio
            lo.append(min(a, b)); hi.append(max(a, b))
        mn[d], mx[d] = max(lo), min(hi)
    r.dof_min, r.dof_max = mn, mx
    return sc


def auto_grasp_targets(robot, speed_factor=1.0):
    """Hand::autoGrasp (static branch): per-DOF target and step"""
    vel = np.asarray(robot.dof_velocity, dtype=np.float64)
    desired = np.where(speed_factor * vel >= 0, robot.dof_max, robot.dof_min).astype(np.float64)
    steps = vel * speed_factor * AUTO_GRASP_TIME_STEP
    return desired, steps
