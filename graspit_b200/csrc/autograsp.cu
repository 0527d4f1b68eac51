// Batched static DOF closing ("autograsp"): Robot::moveDOFToContacts for thousands of hand poses at once.
//
// Reference control flow (src/robot.cpp): moveDOFToContacts 1708-1815 steps every DOF towards its target;
// jumpDOFToContact 1593-1706 asks the DOFs for joint values (getJointValuesFromDOF 1507-1560, RigidDOF / BreakAwayDOF
// accumulateMove, src/dof.cpp:500-514,577-620), moves the links, and while allCollisions reports a hit among the moving
// links bisects between the last free and the colliding joint values on bodyToBodyDistance of the colliding pairs
// (interpolateJoints 1418-1472) until they are closer than Contact::THRESHOLD / 2 without touching; pairs within
// Contact::THRESHOLD (World::findContacts, world.cpp:1682-1707) stop the joints upstream of their links
// (stopJointsFromLink 1474-1489) and the DOFs take the interpolated values (updateVal, dof.cpp:622-648).
//
// B200 mapping: the geometry (joint-level FK, ALL_COLLISIONS pair masks, pair distances) is evaluated for the whole
// batch by the engine's kernels; this file holds the per-pose state machine, one thread per pose.  Every pose walks
// the reference's loop at its own pace: a round of the host loop is  state step -> FK -> collide -> listed distances,
// and a pose consumes the answers of the previous round (both were computed at the joint values it left in jstates).
// Poses never wait for each other: the number of rounds is the longest single-pose sequence, not the sum.
// Link::contactsPreventMotion inside getJointValuesFromDOF is not evaluated: from a contact-free start it cannot fire
// (see include/b2c.h, b2c_autograsp_batch).
#include "b2c_internal.h"

namespace b2c {

namespace {

__device__ inline bool bit(const AgCol<unsigned long long> &m, int i) { return (m[i >> 6] >> (i & 63)) & 1ull; }

// RigidDOF::accumulateMove (dof.cpp:500-514) / BreakAwayDOF::accumulateMove (dof.cpp:577-620) /
// CompliantDOF::accumulateMove (dof.cpp:830-856: every joint moves on its own unless IT is stopped in that direction;
// joint_ratio holds CompliantDOF::getStaticRatio of the joint)
__device__ bool accumulate_move(const AgDesc &D, const AgState &S, int d, double q1, double *jv) {
  const double q = S.q[d];
  if (fabs(q - q1) < 1.0e-5) return false;
  if (D.dof_breakaway[d] == 2) {
    bool movement = false;
    for (int j = 0; j < D.nj; j++) {
      if (D.joint_dof[j] != d) continue;
      double nv = q1 * D.joint_ratio[j], ov = S.jv[j];
      if (nv > ov && (S.stopped[j] & 1)) continue;
      if (nv < ov && (S.stopped[j] & 2)) continue;
      jv[j] = nv;
      movement = true;
    }
    return movement;
  }
  if (!D.dof_breakaway[d]) {
    for (int j = 0; j < D.nj; j++)
      if (D.joint_dof[j] == d && S.stopped[j]) return false;
    for (int j = 0; j < D.nj; j++)
      if (D.joint_dof[j] == d) jv[j] = q1 * D.joint_ratio[j];
    return true;
  }
  for (int j = 0; j < D.nj; j++)
    if (D.joint_dof[j] == d && S.stopped[j] && q1 < q) return false;
  bool movement = false;
  for (int j = 0; j < D.nj; j++) {
    if (D.joint_dof[j] != d) continue;
    if (S.in_break[j] && q1 > S.break_val[j]) continue;
    double nv = q1 * D.joint_ratio[j], ov = S.jv[j];
    if (nv > ov && (S.stopped[j] & 1)) continue;
    if (nv < ov && (S.stopped[j] & 2)) continue;
    jv[j] = nv;
    movement = true;
  }
  return movement;
}

// one pass of moveDOFToContacts' loop up to the point where jumpDOFToContact has moved the links: false = no DOF moves
__device__ bool begin_jump(const AgDesc &D, AgState &S) {
  double nv[AG_MAXD], njv[AG_MAXJ];
  for (int d = 0; d < D.ndof; d++) {
    double v = S.q[d] + S.step[d];
    if (S.step[d] > 0 && v > S.desired[d]) v = S.desired[d];
    else if (S.step[d] < 0 && v < S.desired[d]) v = S.desired[d];
    nv[d] = v;
  }
  for (int j = 0; j < D.nj; j++) { S.initial_jv[j] = S.jv[j]; njv[j] = S.jv[j]; }
  bool moving = false;
  for (int d = 0; d < D.ndof; d++) {
    S.initial_dof[d] = S.q[d];
    if (accumulate_move(D, S, d, nv[d], njv)) { moving = true; S.new_dof[d] = nv[d]; }
    else S.new_dof[d] = S.q[d];
  }
  if (!moving) return false;
  // interest list: links whose last joint is not stopped (robot.cpp:1623-1631) -> active pairs touching one of them
  for (int w = 0; w < D.mask_words; w++) { S.interest[w] = 0; S.late[w] = 0; }
  for (int p = 0; p < D.npairs; p++) {
    int la = D.pair_link_a[p], lb = D.pair_link_b[p];
    bool in = (la >= 0 && !S.stopped[D.link_last_joint[la]]) || (lb >= 0 && !S.stopped[D.link_last_joint[lb]]);
    if (in) S.interest[p >> 6] |= 1ull << (p & 63);
  }
  for (int j = 0; j < D.nj; j++) S.jv[j] = njv[j];
  return true;
}

// Robot::stopJointsFromLink (robot.cpp:1474-1489)
__device__ void stop_joints_from_link(const AgDesc &D, AgState &S, int link) {
  for (int j = D.link_last_joint[link]; j >= D.link_first_joint[link]; j--) {
    if (S.final_jv[j] > S.jv[j]) S.stopped[j] |= 1;
    else if (S.final_jv[j] < S.jv[j]) S.stopped[j] |= 2;
  }
}

// tail of jumpDOFToContact once the links are collision-free: findContacts' distance filter, joint stopping, updateDofVals
__device__ void finish_jump(const AgDesc &D, AgState &S, const float *dist) {
  int ncols = 0;
  for (int p = 0; p < D.npairs; p++) {
    if (!bit(S.late, p)) continue;
    if ((double)dist[p] > D.threshold) continue;
    ncols++;
    if (D.pair_link_a[p] >= 0) stop_joints_from_link(D, S, D.pair_link_a[p]);
    if (D.pair_link_b[p] >= 0) stop_joints_from_link(D, S, D.pair_link_b[p]);
  }
  S.ncols = ncols;
  for (int d = 0; d < D.ndof; d++) {
    double q1 = S.new_dof[d];
    if (D.dof_breakaway[d] == 1) {       // BreakAwayDOF::updateVal (dof.cpp:622-648); Rigid / CompliantDOF::updateVal: q = q1
      for (int j = 0; j < D.nj; j++) {
        if (D.joint_dof[j] != d) continue;
        double jval = S.jv[j] / D.joint_ratio[j];
        if (S.in_break[j]) {
          if (q1 < S.break_val[j] - 1.0e-5) S.in_break[j] = 0;
        } else if (jval < q1 - 1.0e-5) {
          S.in_break[j] = 1;
          S.break_val[j] = jval;
        }
      }
    }
    S.q[d] = q1;
  }
}

__device__ inline void interp_joints(const AgDesc &D, AgState &S) {
  for (int j = 0; j < D.nj; j++) S.jv[j] = S.t * S.final_jv[j] + (1.0 - S.t) * S.initial_jv[j];
}

__global__ void __launch_bounds__(128) autograsp_step_kernel(AgParams p) {
  const int pose = blockIdx.x * blockDim.x + threadIdx.x;
  if (pose >= p.npose) return;
  const AgDesc &D = *p.desc;
  AgState S(p.st, pose);
  double *js = p.jstates + (size_t)pose * (7 + D.nj);
  if (p.init) {
    const float *s = p.states + (size_t)pose * p.stride;
    for (int k = 0; k < 7; k++) js[k] = (double)s[k];
    for (int d = 0; d < D.ndof; d++) S.q[d] = (double)s[7 + d];
    for (int j = 0; j < D.nj; j++) {
      S.stopped[j] = 0; S.in_break[j] = 0; S.break_val[j] = -10.0;
      S.jv[j] = S.q[D.joint_dof[j]] * D.joint_ratio[j];          // forceDOFVals: DOFs reset, joint = ratio * dof
      S.final_jv[j] = S.jv[j];
    }
    for (int d = 0; d < D.ndof; d++) {                             // moveDOFToContacts: step sizes (robot.cpp:1720-1734)
      double des = D.desired[d], st;
      if (D.one_step) st = des - S.q[d];
      else if (D.step[d] != 0.0) st = (des < S.q[d] ? -1.0 : 1.0) * fabs(D.step[d]);
      else { st = 0.0; des = S.q[d]; }
      S.desired[d] = des; S.step[d] = st;
    }
    S.iters = 1; S.status = 0; S.ncols = 0; S.t = 0; S.deltat = 1;
    for (int w = 0; w < D.mask_words; w++) S.late[w] = S.col[w] = S.interest[w] = 0;
    S.phase = begin_jump(D, S) ? AG_CHECK : AG_DONE;
  } else if (S.phase != AG_DONE) {
    const float *dist = p.dist + (size_t)pose * D.npairs;
    bool check = S.phase == AG_CHECK;
    if (S.phase == AG_INTERP) {                                    // one pass of interpolateJoints' loop body
      double mind = 1.0e10;
      bool hit = false;
      for (int q = 0; q < D.npairs && !hit; q++) {
        if (!bit(S.col, q)) continue;
        double d = (double)dist[q];
        mind = d < mind ? d : mind;
        if (mind <= 0) { S.t -= S.deltat; hit = true; }
      }
      bool done = false;
      if (mind > 0) {
        if (mind < 0.5 * D.threshold) done = true;
        else S.t += S.deltat;
      }
      if (done) {
        // lateContacts = colReport; DOF targets interpolated; the loop of jumpDOFToContact checks collisions again
        for (int w = 0; w < D.mask_words; w++) S.late[w] = S.col[w];
        for (int d = 0; d < D.ndof; d++) S.new_dof[d] = S.new_dof[d] * S.t + S.initial_dof[d] * (1.0 - S.t);
        check = true;
      } else if (S.deltat > 1.0e-20 && S.t >= 0) {
        S.deltat *= 0.5;
        interp_joints(D, S);
      } else {
        S.status |= AG_ST_INTERP_ERROR;                            // jumpDOFToContact returns false, DOF values stay
        S.phase = AG_DONE;
      }
    }
    if (check && S.phase != AG_DONE) {
      const unsigned long long *pm = p.pair_mask + (size_t)pose * p.mask_words;
      bool any = false;
      for (int w = 0; w < D.mask_words; w++) {
        S.col[w] = pm[w] & S.interest[w];
        any |= S.col[w] != 0ull;
      }
      if (any) {
        for (int j = 0; j < D.nj; j++) S.final_jv[j] = S.jv[j];    // getJointValues(newJointVals)
        S.t = 0.0; S.deltat = 0.5;                                 // first pass of the loop: deltat *= 0.5, pose at t = 0
        interp_joints(D, S);
        S.phase = AG_INTERP;
      } else {
        finish_jump(D, S, dist);
        S.status |= AG_ST_MOVED;
        if (D.stop_at_contact && S.ncols != 0) { S.status |= AG_ST_STOPPED_AT_CONTACT; S.phase = AG_DONE; }
        else if (S.iters > D.max_iters) { S.status |= AG_ST_FAILSAFE; S.phase = AG_DONE; }
        else {
          S.iters++;
          S.phase = begin_jump(D, S) ? AG_CHECK : AG_DONE;
        }
      }
    }
  }
  for (int w = 0; w < p.mask_words; w++)
    p.pair_enable[(size_t)pose * p.mask_words + w] = S.phase != AG_DONE ? S.interest[w] : 0ull;
  if (S.phase != AG_DONE) {
    for (int j = 0; j < D.nj; j++) js[7 + j] = S.jv[j];
    atomicAdd(&p.counters[0], 1);
    if (S.phase == AG_INTERP) {
      int n = 0;
      for (int w = 0; w < D.mask_words; w++) n += __popcll(S.col[w]);
      int at = atomicAdd(&p.counters[1], n);
      for (int q = 0; q < D.npairs; q++)
        if (bit(S.col, q)) { if (at < p.wcap) p.wlist[at] = pose * D.npairs + q; at++; }
    }
  }
}

} // namespace

cudaError_t launch_autograsp_step(const AgParams &p, cudaStream_t s) {
  if (p.npose <= 0) return cudaSuccess;
  autograsp_step_kernel<<<(p.npose + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace b2c
