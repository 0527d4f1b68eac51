// CUDA kernels of the B200 collision engine (sm_100a).
//
//   fk_kernel        K1  search state -> hand pose -> per-link transforms          (thread / pose)
//   (legality, K3: collide.cu -- broad_kernel thread / (pose, pair), narrow_kernel warp / 32 work items)
//   recheck_kernel   fp64 re-decision of triangle pairs the fp32 SAT could not call
//   point_kernel     K5 stand-alone closest point queries                            (thread / query)
//   dist_kernel      K4 body-body minimum distance                                   (warp / query)
//   energy_kernel    K5+K8 CONTACT_ENERGY, PRESET or LIVE contacts, TMA-staged tree top (thread / pose,contact)
//
// See DESIGN.md for the data layout and the roofline that bounds each kernel.
#include <cstdlib>
#include "b2c_internal.h"
#include "traverse.cuh"

namespace b2c {

// ============================================================================================
// K1: forward kinematics.  One thread per pose, fp64 quaternion algebra in the reference's
// composition order (KinematicChain::fwdKinematics, src/kinematicChain.cpp:543-560;
// DHTransform::getTran, include/graspit/joint.h:126-130; HandObjectState::execute,
// src/EGPlanner/searchState.cpp:515-527; PositionStateAA::getCoreTran, searchStateImpl.cpp:156-168;
// EigenGraspInterface::toDOFSpace, src/eigenGrasp.cpp:638-646; Robot::checkSetDOFVals, robot.h:698-709).
// ============================================================================================

#define FK_MAX_DOF 32

__global__ void __launch_bounds__(128) fk_kernel(FkParams p) {
  int pose = blockIdx.x * blockDim.x + threadIdx.x;
  if (pose >= p.npose) return;
  const float *s = p.states + (size_t)pose * p.stride;
  const FkProgram &g = p.prog;
  double dof[FK_MAX_DOF];
  Qt base;
  const double *jv = nullptr;      // joint-level input (autograsp): joint values replace dof * ratio
  if (p.input_mode == 2) {
    const double *sj = p.jstates + (size_t)pose * p.stride;
    base.w = sj[0]; base.x = sj[1]; base.y = sj[2]; base.z = sj[3];
    base.tx = sj[4]; base.ty = sj[5]; base.tz = sj[6];
    q_normalize(base);
    jv = sj + 7;
  } else if (p.input_mode == 0) {
    base.w = s[0]; base.x = s[1]; base.y = s[2]; base.z = s[3];
    base.tx = s[4]; base.ty = s[5]; base.tz = s[6];
    q_normalize(base);
    for (int d = 0; d < g.ndof_total && d < FK_MAX_DOF; d++) dof[d] = (double)s[7 + d];
  } else {
    // search-state input (HandObjectState::execute, searchState.cpp:515-527): hand transform = mRefTran % coreTran(position
    // variables), posture from the remaining variables.  coreTran per position type: searchStateImpl.cpp:125-135 (COMPLETE),
    // :156-168 (AXIS_ANGLE), :218-252 (ELLIPSOID), :266-274 (APPROACH)
    const FkRobot &r0 = g.robots[0];
    Qt core;
    int npos;
    if (p.input_mode == 1) {
      double tx = s[0], ty = s[1], tz = s[2], th = s[3], ph = s[4], al = s[5];
      double sth, cth, sph, cph;
      sincos(th, &sth, &cth);
      sincos(ph, &sph, &cph);
      Qt tr = q_identity(); tr.tx = tx; tr.ty = ty; tr.tz = tz;
      core = q_compose(tr, q_axis_angle(al, sth * cph, sth * sph, cth));
      core = q_compose(core, r0.approach_inv);
      npos = 6;
    } else if (p.input_mode == 3) {
      const double a = p.pos_params[0], b = p.pos_params[1], c = p.pos_params[2];
      double sb, cb, sg, cg;
      sincos((double)s[0], &sb, &cb);
      sincos((double)s[1], &sg, &cg);
      const double tau = s[2], dist = s[3];
      d3 pt = mk<double>(a * cb * cg, b * cb * sg, c * sb);
      d3 n1 = d_normalized(mk<double>(-a * sb * cg, -b * sb * sg, c * cb));
      d3 n2 = d_normalized(mk<double>(-a * cb * sg, b * cb * cg, 0.0));
      d3 normal = cross(n1, n2);                       // NOT unit length unless a == b; the reference uses it as it is
      d3 ydir = cross(normal, mk<double>(1.0, 0.0, 0.0));
      d3 xdir = cross(ydir, normal);
      d3 cx = d_normalized(xdir), cy = d_normalized(ydir), cz = d_normalized(normal);
      const double m[9] = {cx.x, cy.x, cz.x, cx.y, cy.y, cz.y, cx.z, cy.z, cz.z};      // columns = x, y, normal
      Qt hand = q_from_matrix(m);
      hand.tx = pt.x - dist * normal.x; hand.ty = pt.y - dist * normal.y; hand.tz = pt.z - dist * normal.z;
      Qt zrot = q_identity();                          // Eigen::AngleAxisd(tau, z) -> Quaternion: no small-angle snap here
      sincos(0.5 * tau, &zrot.z, &zrot.w);
      core = q_compose(q_compose(hand, zrot), r0.approach_inv);
      npos = 4;
    } else if (p.input_mode == 4) {
      Qt hand = q_identity(); hand.tz = s[0];
      hand = q_compose(q_compose(q_axis_angle((double)s[2], 0.0, 1.0, 0.0), q_axis_angle((double)s[1], 1.0, 0.0, 0.0)), hand);
      core = q_compose(q_compose(r0.approach, hand), r0.approach_inv);
      npos = 3;
    } else {
      core.w = s[3]; core.x = s[4]; core.y = s[5]; core.z = s[6];
      core.tx = s[0]; core.ty = s[1]; core.tz = s[2];
      q_normalize(core);
      npos = 7;
    }
    base = q_compose(p.ref, core);
    if (p.input_mode == 5) {
      // POSE_DOF (PostureStateDOF::getHandDOF, searchStateImpl.cpp:44-49): the variables are the DOF values
      for (int d = 0; d < g.ndof_total && d < FK_MAX_DOF; d++) dof[d] = d < r0.ndof ? (double)s[npos + d] : 0.0;
    } else {
      // POSE_EIGEN (PostureStateEigen::getHandDOF, :83-95; EigenGraspInterface::toDOFSpace, eigenGrasp.cpp:638-646; checkSetDOFVals)
      const double *eg = g.egdata + r0.eg_ofs;
      const double *origin = eg + r0.neg * r0.ndof, *norm = origin + r0.ndof, *mn = norm + r0.ndof, *mx = mn + r0.ndof;
      for (int d = 0; d < r0.ndof && d < FK_MAX_DOF; d++) {
        double x = 0.0;
        for (int e = 0; e < r0.neg; e++) x += eg[e * r0.ndof + d] * (double)s[npos + e];
        double v = x * norm[d] + origin[d];
        v = v > mx[d] ? mx[d] : (v < mn[d] ? mn[d] : v);
        dof[d] = v;
      }
      for (int d = r0.ndof; d < g.ndof_total && d < FK_MAX_DOF; d++) dof[d] = 0.0;
    }
  }
  Xf *out = p.out + (size_t)pose * g.nslots;
  double *oq = p.out_qt ? p.out_qt + (size_t)pose * g.nslots * 7 : nullptr;
  auto emit = [&](int slot, const Qt &t) {
    if (slot < 0) return;
    q_to_xf(t, reinterpret_cast<double *>(&out[slot]));
    if (oq) {
      double *o = oq + slot * 7;
      o[0] = t.w; o[1] = t.x; o[2] = t.y; o[3] = t.z; o[4] = t.tx; o[5] = t.ty; o[6] = t.tz;
    }
  };
  // last-link transforms of chains that carry an attached robot are re-read from this small table
  Qt lastlink[4];                  // FK_MAX_ROBOTS (build_plan rejects larger trees)
  for (int ri = 0; ri < g.nrobots; ri++) {
    const FkRobot &r = g.robots[ri];
    Qt rb = base;
    if (r.parent >= 0) rb = q_compose(lastlink[ri], r.parent_offset);
    emit(r.base_slot, rb);
    emit(r.mount_slot, rb);
    for (int ci = 0; ci < r.nchains; ci++) {
      const FkChain &c = g.chains[r.first_chain + ci];
      Qt total = q_compose(rb, c.tran);
      int l = 0;
      for (int j = 0; j < c.njoints; j++) {
        const FkJoint &jd = g.joints[c.first_joint + j];
        double val = jv ? jv[c.first_joint + j] : dof[jd.dof] * jd.ratio;
        double theta = jd.revolute ? val + jd.theta : jd.theta;
        double dd = jd.revolute ? jd.d : val + jd.d;
        Qt tz = q_identity(); tz.tz = dd;
        Qt jt = q_compose(q_compose(q_axis_angle(theta, 0.0, 0.0, 1.0), tz), jd.tr4tr3);
        total = q_compose(total, jt);
        if (l < c.nlinks && g.links[c.first_link + l].last_joint == j) {
          emit(g.links[c.first_link + l].out_slot, total);
          l++;
        }
      }
      // children attached to this chain's end: remember the last link pose for them
      for (int rj = ri + 1; rj < g.nrobots; rj++)
        if (g.robots[rj].parent == ri && g.robots[rj].parent_last_slot == g.links[c.first_link + c.nlinks - 1].out_slot)
          lastlink[rj] = total;
    }
  }
}

// ============================================================================================
// closest point on a mesh to a query point (object frame).  Per-lane depth-first, nearer child
// first; boxes cull in fp32 (they are padded at build time), leaves are re-centred on the query
// in fp64 before the fp32 closestPtTriangle so the returned vector carries fp32 *relative* accuracy.
// Must agree with ClosestPtCallback (collisionAlgorithms_inl.h:210-238, collisionAlgorithms.h:220-242).
// ============================================================================================
// out-of-line fp64 helpers: keep their register needs out of the hot fp32 paths
__device__ __noinline__ d3 closest_pt_triangle_d(d3 a, d3 b, d3 c) {
  return closest_pt_triangle<double>(a, b, c, mk<double>(0.0, 0.0, 0.0));
}
__device__ __noinline__ double tri_tri_dist2_d(d3 a1, d3 a2, d3 a3, d3 b1, d3 b2, d3 b3, d3 *pa, d3 *pb) {
  d3 x, y;
  double d = tri_tri_dist2<double>(a1, a2, a3, b1, b2, b3, x, y);
  *pa = x; *pb = y;
  return d;
}

// rare paths of the distance leaf test, kept out of line so the traversal loop stays small in the i-cache
__device__ __noinline__ int tri_tri_filter_call(f3 p2, f3 p3, f3 e1, f3 e2, f3 q1, f3 q2, f3 q3, f3 f1, f3 f2) {
  return tri_tri_filter(p2, p3, e1, e2, q1, q2, q3, f1, f2);
}
__device__ __noinline__ float tri_tri_dist2_plain(f3 p2, f3 p3, f3 q1, f3 q2, f3 q3, f3 *pa, f3 *pb) {
  f3 x, y;
  float d = tri_tri_dist2<float>(mk<float>(0.f, 0.f, 0.f), p2, p3, q1, q2, q3, x, y);
  *pa = x; *pb = y;
  return d;
}

// ============================================================================================
// closest point on a mesh to a query point (object frame).  Per-lane depth-first, nearer child
// first, "while-while" form: a lane descends inner nodes until it holds a leaf, then all lanes of the
// warp that hold a leaf run the triangle test together.  Boxes cull in fp32 (they are padded at
// build time); a leaf is re-centred on the query in fp64 before the fp32 closestPtTriangle, so the
// returned vector carries fp32 *relative* accuracy; when the winning distance is tiny against the
// triangle's extent the winner is redone in fp64.
// Must agree with ClosestPtCallback (collisionAlgorithms_inl.h:210-238, collisionAlgorithms.h:220-242).
// ============================================================================================
// Candidates whose squared distance is within this factor of the best are ordered in fp64, not fp32.  Pruning uses the
// same band, so every such candidate is visited whatever order the traversal takes (deterministic results).
constexpr float NEAR_BAND = 1.0002f;
// The same for the body-body distance, whose near-ties are LISTED for the fp64 pass: fp32 on coordinates re-centred next to
// the query resolves squared distances to ~2e-7 relative at 300 mm; the band is 100 x that.  (With 2e-4 every far query
// against the 220 800-triangle mug met 64+ "ties" -- neighbours of the closest triangle on the same facet -- and the
// lists overflowed.)
constexpr float DIST_BAND = 1.00002f;
#define CPQ_STACK 64
__device__ __forceinline__ float closest_point_query(const MeshDev &m, const NodeQ *top, int ntop, d3 q, f3 &bestv,
                                                     int &best_tri, unsigned &n_box, unsigned &n_tri, int *second_tri = nullptr) {
  const f3 qf = tof(q);
  float best = 1.0e30f, best_sc = 0.f;
  // runner-up with a different closest point inside the near-tie band (arbitrated in fp64 by the caller)
  float second = 3.0e38f; int tri2 = -1; f3 v2 = mk<float>(0.f, 0.f, 0.f);
  bestv = mk<float>(0.f, 0.f, 0.f);
  best_tri = -1;
  int stk_n[CPQ_STACK];
  float stk_d[CPQ_STACK];
  int sp = 0;
  int cur = 0;          // node to visit, -1 = none
  float curd = 0.f;
  while (true) {
    int leaf = -1;
    // ---- descend until this lane holds a leaf or runs out of work
    while (cur >= 0) {
      if (curd <= best * NEAR_BAND + 1.0e-12f) {
        Node n = fetch_node(m, top, ntop, cur);
        if (n.child < 0) {
          leaf = n.tri;
          cur = -1;
        } else {
          Node c0, c1;
          fetch_pair(m, top, ntop, n.child, c0, c1);
          float cc[3], hh[3];
          node_ch(c0, cc, hh); float d0 = point_box_dist2(cc, hh, qf);
          node_ch(c1, cc, hh); float d1 = point_box_dist2(cc, hh, qf);
          n_box += 2;
          int near = n.child, far = n.child + 1;
          if (d1 < d0) { near = n.child + 1; far = n.child; float t = d0; d0 = d1; d1 = t; }
          if (d1 <= best * NEAR_BAND + 1.0e-12f && sp < CPQ_STACK) { stk_n[sp] = far; stk_d[sp] = d1; sp++; }
          if (d0 <= best * NEAR_BAND + 1.0e-12f) { cur = near; curd = d0; continue; }
          cur = -1;
        }
      } else {
        cur = -1;
      }
      if (cur < 0 && leaf < 0 && sp > 0) { sp--; cur = stk_n[sp]; curd = stk_d[sp]; }
    }
    // ---- triangle test for every lane that holds a leaf
    if (leaf >= 0) {
      n_tri++;
      d3 a = tri_vertex(m.tris, leaf, 0) - q, b = tri_vertex(m.tris, leaf, 1) - q, c = tri_vertex(m.tris, leaf, 2) - q;
      f3 af = tof(a), bf = tof(b), cf = tof(c);
      f3 v = closest_pt_triangle<float>(af, bf, cf, mk<float>(0.f, 0.f, 0.f));
      float d2 = dot(v, v);
      auto differs = [](f3 x, f3 y) { return fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) > 2.0e-4f; };
      if (d2 < best) {
        // the displaced best becomes the runner-up if it describes another point and is still inside the band
        if (best_tri >= 0 && best <= d2 * NEAR_BAND + 1.0e-12f && differs(bestv, v)) { second = best; tri2 = best_tri; v2 = bestv; }
        else if (tri2 >= 0 && (second > d2 * NEAR_BAND + 1.0e-12f || !differs(v2, v))) { tri2 = -1; second = 3.0e38f; }
        best = d2; bestv = v; best_tri = leaf;
        best_sc = fmaxf(fmaxf(fmaxf(fabsf(af.x), fabsf(af.y)), fmaxf(fabsf(af.z), fabsf(bf.x))),
                        fmaxf(fmaxf(fabsf(bf.y), fabsf(bf.z)), fmaxf(fabsf(cf.x), fmaxf(fabsf(cf.y), fabsf(cf.z)))));
      } else if (d2 <= best * NEAR_BAND + 1.0e-12f && d2 < second && differs(v, bestv)) {
        second = d2; tri2 = leaf; v2 = v;
      }
      if (sp > 0) { sp--; cur = stk_n[sp]; curd = stk_d[sp]; }
    }
    if (cur < 0) break;
  }
  if (second_tri) *second_tri = tri2;
  // fp32 loses relative accuracy once the distance is small against the triangle's own extent
  if (best_tri >= 0 && best < 1.0e-4f * best_sc * best_sc) {
    d3 a = tri_vertex(m.tris, best_tri, 0) - q, b = tri_vertex(m.tris, best_tri, 1) - q, c = tri_vertex(m.tris, best_tri, 2) - q;
    d3 v64 = closest_pt_triangle_d(a, b, c);
    bestv = tof(v64);
    best = (float)dot(v64, v64);
  }
  return best;
}

// fp64 re-run of a closest-point query whose fp32 traversal met near-ties: same tree, every candidate inside
// `bound2` (squared) evaluated in double, the true minimum kept.  Rare path of the single-pose API (point_kernel).
__device__ __noinline__ double closest_point_query_d(const MeshDev &m, d3 q, double bound2, d3 *closest) {
  int stk[CPQ_STACK];
  int sp = 0;
  stk[sp++] = 0;
  double best = bound2;
  d3 bestc = q;
  bool found = false;
  while (sp > 0) {
    const int i = stk[--sp];
    Node n = load_node(m, i);
    // stored boxes are padded: the fp32 box distance (minus a rounding margin) is a valid lower bound
    float c[3], h[3];
    node_ch(n, c, h);
    double dx = fmax(fabs(q.x - (double)c[0]) - (double)h[0], 0.0), dy = fmax(fabs(q.y - (double)c[1]) - (double)h[1], 0.0),
           dz = fmax(fabs(q.z - (double)c[2]) - (double)h[2], 0.0);
    if (dx * dx + dy * dy + dz * dz > best) continue;
    if (n.child < 0) {
      d3 a = tri_vertex(m.tris, n.tri, 0), b = tri_vertex(m.tris, n.tri, 1), cc = tri_vertex(m.tris, n.tri, 2);
      d3 cl = closest_pt_triangle<double>(a, b, cc, q);
      d3 dv = cl - q;
      double d2 = dot(dv, dv);
      if (d2 < best || (!found && d2 <= best)) { best = d2; bestc = cl; found = true; }
    } else if (sp + 2 <= CPQ_STACK) {
      stk[sp++] = n.child; stk[sp++] = n.child + 1;
    }
  }
  *closest = bestc;
  return found ? best : -1.0;
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ContactEnergy::energy (src/EGPlanner/energy/contactEnergy.cpp:9-55) for one contact: world contact
// location p and world normal cn -> |v| + 50 (1 - cn . v/|v|), v = closest object point - p.
__device__ __forceinline__ float contact_term(const MeshDev &mo, const NodeQ *top, int ntop, const double *XO, d3 pw, d3 cnw,
                                              unsigned &n_box, unsigned &n_tri) {
  d3 q = xf_apply_inv(XO, pw);
  f3 cnl = tof(xf_rot_t(XO, cnw));
  f3 v; int bt;
  float d2 = closest_point_query(mo, top, ntop, q, v, bt, n_box, n_tri);
  float dist = sqrtf(d2);
  float inv = d2 > 0.f ? 1.0f / dist : 0.f;
  float c = (cnl.x * v.x + cnl.y * v.y + cnl.z * v.z) * inv;
  return dist + (1.0f - c) * 50.0f;
}

// ============================================================================================
// fp64 re-decision of ambiguous triangle pairs (reference arithmetic, see tri_tri_intersect_exact).
// A confirmed intersection clears legal/energy, sets the pair bit, and (distance queries) forces -1.
// ============================================================================================

__global__ void recheck_kernel(RecheckParams p) {
  int n = *p.ambig_count;
  if (n > p.ambig_cap) n = p.ambig_cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Ambig a = p.ambig[i];
    int2 ab = p.pairs[a.pair];
    const double *XA = body_xf_ptr(p.fk, p.static_xf, p.nrb, a.pose, ab.x);
    const double *XB = body_xf_ptr(p.fk, p.static_xf, p.nrb, a.pose, ab.y);
    const MeshDev &ma = p.meshes[p.body_mesh[ab.x]];
    const MeshDev &mb = p.meshes[p.body_mesh[ab.y]];
    d3 a1 = xf_apply_inv(XB, xf_apply(XA, tri_vertex(ma.tris, a.tri_a, 0)));
    d3 a2 = xf_apply_inv(XB, xf_apply(XA, tri_vertex(ma.tris, a.tri_a, 1)));
    d3 a3 = xf_apply_inv(XB, xf_apply(XA, tri_vertex(ma.tris, a.tri_a, 2)));
    d3 b1 = tri_vertex(mb.tris, a.tri_b, 0), b2 = tri_vertex(mb.tris, a.tri_b, 1), b3 = tri_vertex(mb.tris, a.tri_b, 2);
    if (tri_tri_intersect_exact(a1, a2, a3, b1, b2, b3)) {
      if (p.legal) p.legal[a.pose] = 0;
      if (p.energy) p.energy[a.pose] = 0.f;
      if (p.pair_mask) atomicOr(&p.pair_mask[(size_t)a.pose * p.mask_words + (a.pair >> 6)], 1ull << (a.pair & 63));
      if (p.dist) {
        p.dist[(size_t)a.pose * p.nq + a.pair] = -1.f;
        if (p.pts) for (int k = 0; k < 6; k++) p.pts[((size_t)a.pose * p.nq + a.pair) * 6 + k] = 0.0;
      }
      if (p.stats) atomicAdd(&p.stats[STAT_HITFIX], 1ull);
    }
  }
}

// ============================================================================================
// K5 stand-alone: closest point from world points to one body.  One thread per query.
// pointToBodyDistance (graspitCollision.cpp:435-456).
// ============================================================================================

__global__ void point_kernel(PointParams p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const double *X = reinterpret_cast<const double *>(p.body_xf);
  d3 pw = mk<double>(p.points[3 * i], p.points[3 * i + 1], p.points[3 * i + 2]);
  d3 q = xf_apply_inv(X, pw);
  f3 v; int bt, bt2 = -1; unsigned nb = 0, nt = 0;
  float d2 = closest_point_query(p.mesh, nullptr, 0, q, v, bt, nb, nt, &bt2);
  // refine the winning triangle in fp64 so the returned coordinates carry fp64 accuracy; a runner-up with another
  // closest point inside the near-tie band is compared in fp64 (the reference's fp64 traversal would have ordered them)
  d3 cl = q;
  double dist = 0.0;
  if (bt >= 0) {
    d3 a = tri_vertex(p.mesh.tris, bt, 0), b = tri_vertex(p.mesh.tris, bt, 1), c = tri_vertex(p.mesh.tris, bt, 2);
    cl = closest_pt_triangle<double>(a, b, c, q);
    d3 dv = cl - q;
    double dd = dot(dv, dv);
    if (bt2 >= 0) {
      d3 cl2;
      double d2r = closest_point_query_d(p.mesh, q, dd * (1.0 + 1.0e-9), &cl2);
      if (d2r >= 0.0 && d2r < dd) { cl = cl2; dd = d2r; }
    }
    dist = sqrt(dd);
  }
  (void)d2;
  d3 cw = xf_apply(X, cl);
  d3 nv = cw - pw;
  double nn = sqrt(dot(nv, nv));
  if (nn > 0.0) nv = nv * (1.0 / nn);
  double *o = p.out + 7 * i;
  o[0] = dist; o[1] = cw.x; o[2] = cw.y; o[3] = cw.z; o[4] = nv.x; o[5] = nv.y; o[6] = nv.z;
}

// ============================================================================================
// K4: body-body minimum distance.  One warp per (pose, query).  Same cooperative work stack as the
// collision traversal, but every entry carries a lower bound and the warp shares the best distance
// found so far: entries whose bound exceeds it are dropped when popped.  Children are pushed far
// first, near last, so the top of the stack is the most promising region (the reference descends
// the nearer child first, collisionAlgorithms.cpp:98-101).
// Agrees with DistanceCallback (collisionAlgorithms_inl.h:168-204, collisionAlgorithms.h:187-210):
// -1 as soon as any triangle pair intersects, else the minimum over triangleTriangleDistanceSq.
// ============================================================================================
__device__ __forceinline__ unsigned long long pack_dist_entry(float lb, int a, int b) {
  // 16-bit truncated (rounded toward zero => still a lower bound) float | nodeA:24 | nodeB:24
  unsigned hb = __float_as_uint(lb) >> 16;
  return ((unsigned long long)hb << 48) | ((unsigned long long)(unsigned)a << 24) | (unsigned long long)(unsigned)b;
}
__device__ __forceinline__ float entry_lb(unsigned long long e) { return __uint_as_float(((unsigned)(e >> 48)) << 16); }


constexpr int NEAR_TINY_BIT = 1 << 30;   // marks near-list entries kept because their fp32 distance is unreliable
// ---- near-minimum list of the distance kernel (rare paths, kept out of line: the traversal loop is register-bound) ----
// drop entries the improving best has left behind; returns the new length
__device__ __noinline__ int near_compact(int2 *nearq, float *neard, float *nearf, int nnear, float best) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  int kept = 0;
  for (int b0 = 0; b0 < nnear; b0 += 32) {
    const int i = b0 + lane;
    bool k = false; int2 en = make_int2(0, 0); float dn = 0.f, fn = 0.f;
    if (i < nnear) { en = nearq[i]; dn = neard[i]; fn = nearf[i]; k = dn <= best * ((en.y & NEAR_TINY_BIT) ? 1.004f : DIST_BAND) + 1.0e-9f; }
    const unsigned km = __ballot_sync(FULL, k);
    __syncwarp();
    if (k) { const int at = kept + __popc(km & lt_mask); nearq[at] = en; neard[at] = dn; nearf[at] = fn; }
    kept += __popc(km);
    __syncwarp();
  }
  return kept;
}
// Pairs still within tolerance of the FINAL best go to the fp64 refinement kernel.  All of them when there is a decision
// to make: two or more candidates with different closest points, or a distance fp32 could not resolve.  Candidates that
// tie because they share the closest vertex or edge describe one solution: then only the winner goes, so that the
// closest POINTS reported for a query always come from fp64 arithmetic -- between nearly parallel features the location of
// the closest points is ill-conditioned (fp32 on coordinates of a few hundred mm moved them by 0.1 mm while the distance
// changed by 2e-8 relative), and CONTACT_LIVE builds its virtual contacts from those points.
__device__ __noinline__ void near_emit(const DistParams &p, const int2 *nearq, const float *neard, const float *nearf, int nnear,
                                       float best, float best_fp, bool overflow, int w) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  __syncwarp();
  int cnt = 0; bool anytiny = false, anydiff = false;
  unsigned km[(DNEAR_CAP + 31) / 32];
#pragma unroll
  for (int r = 0; r < (DNEAR_CAP + 31) / 32; r++) {
    const int i = r * 32 + lane;
    bool k = false, t = false, differs = false;
    if (i < nnear) {
      t = (nearq[i].y & NEAR_TINY_BIT) != 0;
      k = neard[i] <= best * (t ? 1.004f : DIST_BAND) + 1.0e-9f;
      differs = k && __float_as_int(nearf[i]) != __float_as_int(best_fp);
    }
    km[r] = __ballot_sync(FULL, k);
    anytiny = anytiny || __any_sync(FULL, k && t);
    anydiff = anydiff || __any_sync(FULL, differs);
    cnt += __popc(km[r]);
  }
  if (cnt == 0) return;
  const bool arbitrate = (cnt >= 2 && anydiff) || anytiny;
  if (!arbitrate && !p.pts) return;          // distances only (autograsp work lists): fp32 is what the caller asked for
  if (!arbitrate && cnt > 1) {
    // no arbitration needed: keep the fp32 winner alone (smallest distance, lowest index on ties)
    float bd = 3.0e38f; int bi = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < (DNEAR_CAP + 31) / 32; r++) {
      const int i = r * 32 + lane;
      if (((km[r] >> lane) & 1u) && (neard[i] < bd)) { bd = neard[i]; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(FULL, bd, o); const int oi = __shfl_xor_sync(FULL, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
#pragma unroll
    for (int r = 0; r < (DNEAR_CAP + 31) / 32; r++) km[r] = (bi >> 5) == r ? (1u << (bi & 31)) : 0u;
    cnt = 1;
  }
  int seg = 0, ofs = 0;
  if (lane == 0) { seg = atomicAdd(p.near_count, 1); ofs = atomicAdd(p.near_count + 1, cnt); }
  seg = __shfl_sync(FULL, seg, 0); ofs = __shfl_sync(FULL, ofs, 0);
  if (seg >= p.near_cap_dir || ofs + cnt > p.near_cap_ent) {
    // no room for this query's fp64 pass: its fp32 answer stands, and the caller hears about it
    if (p.overflow && lane == 0) atomicOr(p.overflow, OVF_DIST_QUEUE);
    return;
  }
  int at0 = 0;
#pragma unroll
  for (int r = 0; r < (DNEAR_CAP + 31) / 32; r++) {
    const int i = r * 32 + lane;
    if ((km[r] >> lane) & 1u) { int2 en = nearq[i]; en.y &= ~NEAR_TINY_BIT; p.near_ent[ofs + at0 + __popc(km[r] & lt_mask)] = en; }
    at0 += __popc(km[r]);
  }
  if (lane == 0) p.near_dir[seg] = make_int4(w, ofs, cnt, overflow ? 0 : 1);
}

#ifndef DLEAF_TRIGGER
#define DLEAF_TRIGGER 32     // leaf pairs queued before a batch of triangle tests runs
#endif
#ifndef DIST_MIN_BLOCKS
#define DIST_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(256, DIST_MIN_BLOCKS) dist_kernel(DistParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned char *base = smem_raw + (size_t)warp * DIST_WARP_SMEM;
  unsigned long long *stack = reinterpret_cast<unsigned long long *>(base);
  unsigned long long *leafq = stack + DSTACK_PHYS;
  double *XA = reinterpret_cast<double *>(leafq + DLEAF_CAP);
  double *XB = XA + 12;
  double *XR = XB + 12;                      // B <- A, fp64: R = Rb^T Ra (row-major), t = Rb^T (ta - tb)
  PairXf *xp = reinterpret_cast<PairXf *>(XR + 12);
  int2 *nearq = reinterpret_cast<int2 *>(xp + 1);
  unsigned long long *confq = reinterpret_cast<unsigned long long *>(nearq + DNEAR_CAP);   // [DCONF_CAP] pairs that passed the pre-filter
  float *neard = reinterpret_cast<float *>(confq + DCONF_CAP);                             // [DNEAR_CAP] fp32 squared distance of nearq entries
  float *nearf = neard + DNEAR_CAP;                                                          // [DNEAR_CAP] fingerprint (hash bits) of their closest points
  unsigned n_box = 0, n_tri = 0;
#ifdef DIST_DEBUG_STATS
  unsigned dbg0 = 0, dbg1 = 0, dbg2 = 0, dbg3 = 0;
#endif
  const int total = p.wlist ? min(*p.wcount, p.wcap) : p.npose * p.nq;

  while (true) {
    int w = 0;
    if (lane == 0) {
      w = atomicAdd(p.work_counter, 1);
      // optional compact work list (autograsp bisection): entry = pose * nq + query; other entries keep their value
      if (p.wlist) w = w < total ? __ldg(&p.wlist[w]) : 0x7fffffff;
    }
    w = __shfl_sync(FULL, w, 0);
    if (w >= (p.wlist ? 0x7fffffff : total)) break;
    int pose = w / p.nq, qi = w - pose * p.nq;
    if (p.legal && !p.legal[pose]) {
      if (lane == 0) { p.dist[w] = -1.f; if (p.pts) for (int k = 0; k < 6; k++) p.pts[(size_t)w * 6 + k] = 0.0; }
      continue;
    }
    int2 ab = __ldg(&p.queries[qi]);
    const MeshDev &ma = p.meshes[__ldg(&p.body_mesh[ab.x])];
    const MeshDev &mb = p.meshes[__ldg(&p.body_mesh[ab.y])];
    {
      const double *sa = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.x);
      const double *sb = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.y);
      if (lane < 12) XA[lane] = sa[lane];
      else if (lane < 24) XB[lane - 12] = sb[lane - 12];
    }
    __syncwarp();
    if (lane == 0) { PairXf x; make_pair_xf(XA, XB, ma.extent, mb.extent, x); *xp = x; }
    else if (lane >= 8 && lane < 17) {
      int i = (lane - 8) / 3, j = (lane - 8) % 3;
      XR[3 * i + j] = XB[i] * XA[j] + XB[3 + i] * XA[3 + j] + XB[6 + i] * XA[6 + j];
    } else if (lane >= 20 && lane < 23) {
      int i = lane - 20;
      XR[9 + i] = XB[i] * (XA[9] - XB[9]) + XB[3 + i] * (XA[10] - XB[10]) + XB[6 + i] * (XA[11] - XB[11]);
    }
    __syncwarp();
    const PairXf x = *xp;
    // fp32 copy of B <- A for the pre-filter (its rounding is covered by the filter's margin)
    float rr[12];
#pragma unroll
    for (int i = 0; i < 12; i++) rr[i] = (float)XR[i];

    float best = 3.0e38f;          // squared distance, warp-uniform
    bool seeded = false;           // false until the first leaf batch: greedy nearest-child descent for a first bound
    d3 bpa = mk<double>(0, 0, 0), bpb = mk<double>(0, 0, 0);   // valid on every lane after each leaf batch
    f3 udir = mk<float>(0.f, 0.f, 0.f);   // unit vector from the best point on A to the best point on B (B's frame)
    f3 udirA = udir;                      // the same direction in A's frame, and its product with the A <- B translation
    float udt = 0.f;
    bool hit = false;
    bool near_overflow = false;    // the near-minimum list lost entries: its fp64 minimum is no longer authoritative
    float best_fp = 0.f;           // fingerprint (hash bits carried in a float) of the best pair's closest points
    int sp = 0, lq = 0, cq = 0, nnear = 0;
    if (lane == 0) stack[0] = pack_dist_entry(0.f, 0, 0);
    sp = 1;
    __syncwarp();
#ifdef DIST_ORACLE_SEED
    // experiment: start from the answer of a previous run (how much of the work is unavoidable with these bounds?)
    if (p.dist[w] > 0.f && p.pts) {
      float dprev = p.dist[w] * 1.0001f;
      best = dprev * dprev; seeded = true;
      const double *o = p.pts + (size_t)w * 6;
      d3 pa0 = xf_apply(XR, mk<double>(o[0], o[1], o[2]));
      bpa = pa0; bpb = mk<double>(o[3], o[4], o[5]);
      float ux = (float)(bpb.x - bpa.x), uy = (float)(bpb.y - bpa.y), uz = (float)(bpb.z - bpa.z);
      float inv = rsqrtf(fmaxf(ux * ux + uy * uy + uz * uz, 1.0e-30f));
      udir = mk<float>(ux * inv, uy * inv, uz * inv);
      udirA = mk<float>(x.r[0] * udir.x + x.r[1] * udir.y + x.r[2] * udir.z, x.r[3] * udir.x + x.r[4] * udir.y + x.r[5] * udir.z,
                        x.r[6] * udir.x + x.r[7] * udir.y + x.r[8] * udir.z);
      udt = udirA.x * x.t[0] + udirA.y * x.t[1] + udirA.z * x.t[2];
    }
#endif

    while (true) {
      if (lq >= DLEAF_TRIGGER || (lq > 0 && (sp == 0 || !seeded))) {
        // ---- pre-filter: up to 32 queued leaf pairs against the separating direction of the best pair so far.
        // For any unit u, dist(tA, tB) >= min_k u.b_k - max_k u.a_k; with u taken from the current closest points the
        // bound is tight exactly where the box bound is loose (far apart, nearly facing surfaces), ~80 instructions
        // instead of a full triangle-triangle distance.  Survivors move to confq.
        int n = lq < 32 ? lq : 32;
        bool keep = false;
        unsigned long long e = 0ull;
        if (lane < n) {
          e = leafq[lq - n + lane];
          keep = entry_lb(e) <= best * DIST_BAND + 1.0e-12f;
          if (keep && seeded) {
            int ta = (int)((e >> 24) & 0xffffffu), tb = (int)(e & 0xffffffu);
            float amax = -3.0e38f, bmin = 3.0e38f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
              float4 va = __ldg(&ma.tris[ta].v[k]), vb = __ldg(&mb.tris[tb].v[k]);
              float ax = rr[0] * va.x + rr[1] * va.y + rr[2] * va.z + rr[9];
              float ay = rr[3] * va.x + rr[4] * va.y + rr[5] * va.z + rr[10];
              float az = rr[6] * va.x + rr[7] * va.y + rr[8] * va.z + rr[11];
              amax = fmaxf(amax, udir.x * ax + udir.y * ay + udir.z * az);
              bmin = fminf(bmin, udir.x * vb.x + udir.y * vb.y + udir.z * vb.z);
            }
            float g = bmin - amax - 4.0f * x.eps;
            keep = !(g > 0.f && g * g > best * 1.01f);
          }
        }
        lq -= n;
        unsigned km = __ballot_sync(FULL, keep);
        if (keep) confq[cq + __popc(km & lt_mask)] = e;
        cq += __popc(km);
        __syncwarp();
        if (!(cq >= 32 || (cq > 0 && (!seeded || (sp == 0 && lq == 0))))) continue;
      }
      if (cq >= 32 || (cq > 0 && (!seeded || (sp == 0 && lq == 0)))) {
        // ---- full test of up to 32 surviving pairs
        seeded = true;
        int n = cq < 32 ? cq : 32;
        float d2 = 3.0e38f;
        int res = TT_SEP;
        int ta = 0, tb = 0;
        d3 P1 = mk<double>(0, 0, 0);
        f3 pa = mk<float>(0, 0, 0), pb = pa;
        bool tiny = false;
        if (lane < n) {
          unsigned long long e = confq[cq - n + lane];
          ta = (int)((e >> 24) & 0xffffffu); tb = (int)(e & 0xffffffu);
          if (entry_lb(e) <= best * DIST_BAND + 1.0e-12f) {
            n_tri++;
            d3 a1 = tri_vertex(ma.tris, ta, 0), a2 = tri_vertex(ma.tris, ta, 1), a3 = tri_vertex(ma.tris, ta, 2);
            d3 b1 = tri_vertex(mb.tris, tb, 0), b2 = tri_vertex(mb.tris, tb, 1), b3 = tri_vertex(mb.tris, tb, 2);
            P1 = xf_apply(XR, a1);
            d3 E1 = xf_rot(XR, a2 - a1);
            d3 E2 = xf_rot(XR, a3 - a2);
            f3 e1 = tof(E1), e2 = tof(E2);
            f3 p2 = e1, p3 = tof(E1 + E2);
            f3 q1 = tof(b1 - P1), q2 = tof(b2 - P1), q3 = tof(b3 - P1);
            // boxes that are certainly apart cannot hold intersecting triangles: skip the SAT
            res = entry_lb(e) > 0.f ? TT_SEP : tri_tri_filter_call(p2, p3, e1, e2, q1, q2, q3, tof(b2 - b1), tof(b3 - b2));
            if (res != TT_HIT) {
              bool degenerate;
              d2 = tri_tri_dist2_compact<float>(mk<float>(0, 0, 0), p2, p3, q1, q2, q3, pa, pb, degenerate, best * 1.01f + 1.0e-9f);
              if (degenerate) d2 = tri_tri_dist2_plain(p2, p3, q1, q2, q3, &pa, &pb);
#ifdef DIST_DEBUG_STATS
              if (d2 > 2.9e38f) dbg0++;                 // plane-bound early outs
              else if (d2 < best) dbg1++;               // tests that beat the batch-start best
              else if (d2 > best * 1.21f) dbg2++;       // tests more than 10% (in distance) beyond best
              if (sqrtf(best) > 50.f) dbg3++;           // tests done while best > 50 mm
#endif
              // distance small against the triangles' extent: fp32 has lost relative accuracy; remember the
              // pair (if it is within rounding of the best so far) for the fp64 pass after the traversal
              float sc = fmaxf(fmaxf(fmaxf(fabsf(p2.x), fabsf(p2.y)), fmaxf(fabsf(p2.z), fabsf(p3.x))), fmaxf(fabsf(p3.y), fabsf(p3.z)));
              sc = fmaxf(sc, fmaxf(fmaxf(fabsf(q1.x), fabsf(q1.y)), fabsf(q1.z)));
              sc = fmaxf(sc, fmaxf(fmaxf(fabsf(q2.x), fabsf(q2.y)), fabsf(q2.z)));
              sc = fmaxf(sc, fmaxf(fmaxf(fabsf(q3.x), fabsf(q3.y)), fabsf(q3.z)));
              tiny = d2 < 1.0e-4f * sc * sc && d2 <= best * 1.004f + 1.0e-9f;
            }
          }
        }
        cq -= n;
        float fpr_lane = 0.f;
        {
          // near-minimum list: every pair within 1e-4 (relative, in distance) of the best so far, plus the pairs whose
          // fp32 distance is unreliable (tiny against their extent).  fp32 cannot order candidates that close, and the
          // reference's outputs depend on WHICH pair wins (closest points feed CONTACT_LIVE), so the final choice among
          // them is made in fp64 by dist_refine_kernel.  `best` here is the value before this batch: a superset.
          // fingerprint of this pair's closest points (B's frame): pairs that share the closest vertex / edge tie exactly
          // but describe the SAME solution and need no arbitration
          // (a hash of both closest points on a 1/512 mm grid in B's frame, carried as the bits of a float: two pairs hash
          // alike only if they describe the same two points; a linear form of the coordinates, used before, let pairs
          // 0.1 mm apart pass as one solution at 300 mm coordinates and left 4 of 43 000 LIVE energies 1e-4 .. 1.6e-3 off)
          unsigned hsh = 2166136261u;
          {
            const double qx[6] = {P1.x + (double)pa.x, P1.y + (double)pa.y, P1.z + (double)pa.z, P1.x + (double)pb.x, P1.y + (double)pb.y, P1.z + (double)pb.z};
#pragma unroll
            for (int k6 = 0; k6 < 6; k6++) hsh = (hsh ^ (unsigned)__double2int_rd(qx[k6] * 512.0)) * 16777619u;
          }
          const float fpr = fpr_lane = __int_as_float((int)hsh);
          const bool cand = res != TT_HIT && d2 < 2.9e38f && (tiny || d2 <= best * DIST_BAND + 1.0e-12f);
          const unsigned cm = __ballot_sync(FULL, cand);
          const int add = __popc(cm);
          if (add) {
            if (nnear + add > DNEAR_CAP) nnear = near_compact(nearq, neard, nearf, nnear, best);
            const int at = nnear + __popc(cm & lt_mask);
            if (cand && at < DNEAR_CAP) { nearq[at] = make_int2(ta, tb | (tiny ? NEAR_TINY_BIT : 0)); neard[at] = d2; nearf[at] = fpr; }
            if (nnear + add > DNEAR_CAP) near_overflow = true;
            nnear = nnear + add < DNEAR_CAP ? nnear + add : DNEAR_CAP;
          }
        }
        if (res == TT_AMBIG) {
          int slot = atomicAdd(p.ambig_count, 1);
          if (slot < p.ambig_cap) { Ambig a; a.pose = pose; a.pair = qi; a.tri_a = ta; a.tri_b = tb; p.ambig[slot] = a; }
          else { res = TT_HIT; if (p.overflow) atomicOr(p.overflow, OVF_DIST_QUEUE); }     // fail safe: undecided = interpenetrating
        }
        if (__any_sync(FULL, res == TT_HIT)) { hit = true; break; }
        // warp arg-min
        float m = d2;
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
        if (m < best) {
          unsigned who = __ballot_sync(FULL, d2 == m);
          int src = __ffs(who) - 1;
          best = m;
          // closest points in B's frame, fp64: re-centred fp32 result + fp64 origin
          d3 PA = mk<double>(P1.x + (double)pa.x, P1.y + (double)pa.y, P1.z + (double)pa.z);
          d3 PB = mk<double>(P1.x + (double)pb.x, P1.y + (double)pb.y, P1.z + (double)pb.z);
          bpa.x = __shfl_sync(FULL, PA.x, src); bpa.y = __shfl_sync(FULL, PA.y, src); bpa.z = __shfl_sync(FULL, PA.z, src);
          bpb.x = __shfl_sync(FULL, PB.x, src); bpb.y = __shfl_sync(FULL, PB.y, src); bpb.z = __shfl_sync(FULL, PB.z, src);
          best_fp = __shfl_sync(FULL, fpr_lane, src);
          float ux = (float)(bpb.x - bpa.x), uy = (float)(bpb.y - bpa.y), uz = (float)(bpb.z - bpa.z);
          float inv = rsqrtf(fmaxf(ux * ux + uy * uy + uz * uz, 1.0e-30f));
          udir = mk<float>(ux * inv, uy * inv, uz * inv);
          udirA = mk<float>(x.r[0] * udir.x + x.r[1] * udir.y + x.r[2] * udir.z, x.r[3] * udir.x + x.r[4] * udir.y + x.r[5] * udir.z,
                            x.r[6] * udir.x + x.r[7] * udir.y + x.r[8] * udir.z);
          udt = udirA.x * x.t[0] + udirA.y * x.t[1] + udirA.z * x.t[2];
        }
        __syncwarp();
        continue;
      }
      if (sp == 0) break;
      int room = DSTACK_CAP - sp;
      if (room <= -60) {
        // about to outgrow the physical stack: answer "interpenetrating" (fail safe) and report it
        if (p.overflow && lane == 0) atomicOr(p.overflow, OVF_STACK);
        hit = true;
        break;
      }
      int k = sp < 32 ? sp : 32;
      if (k > room) k = room > 1 ? room : 1;
      if (!seeded) k = 1;          // greedy descent: only the nearest open entry, until a first distance is known
      bool have = lane < k;
      unsigned long long e = have ? stack[sp - 1 - lane] : 0ull;
      sp -= k;
      __syncwarp();
      bool pn = false, pf = false, ln = false, lf = false;      // near / far child: push?, leaf pair?
      unsigned long long cn = 0ull, cf = 0ull;
      if (have && entry_lb(e) <= best * DIST_BAND + 1.0e-12f) {
        int na = (int)((e >> 24) & 0xffffffu), nb = (int)(e & 0xffffffu);
        Node a = load_node(ma, na), b = load_node(mb, nb);
        bool la = a.child < 0, lb = b.child < 0;
        if (la && lb) {
          pn = ln = true;
          cn = pack_dist_entry(entry_lb(e), a.tri, b.tri);
        } else {
          const bool splitA = lb || (!la && (a.hx * a.hy * a.hz > b.hx * b.hy * b.hz));
          // the two children of the split node against the other node: ONE copy of the box test in the loop body
          const MeshDev &cm = splitA ? ma : mb;
          const int cbase = splitA ? a.child : b.child;
          Node chp[2];
          load_pair(cm, cbase, chp[0], chp[1]);
          const Node other = splitA ? b : a;
          const bool oleaf = splitA ? lb : la;
          float d0 = 0.f, d1 = 0.f;
          int i0a = 0, i0b = 0, i1a = 0, i1b = 0;
          bool lf0 = false, lf1 = false;
#pragma unroll 1
          for (int c = 0; c < 2; c++) {
            const Node ch = c == 0 ? chp[0] : chp[1];
            float ca[3], ha[3], cb[3], hb[3];
            if (splitA) { node_ch(ch, ca, ha); node_ch(other, cb, hb); } else { node_ch(other, ca, ha); node_ch(ch, cb, hb); }
            float dc = box_dist_lb(ca, ha, cb, hb, x);
            {
              // gap of the two boxes along the separating direction of the best pair so far (zero vector until seeded):
              // tight where the face-axis bound is loose (distant, nearly facing surfaces)
              float g = (udir.x * cb[0] + udir.y * cb[1] + udir.z * cb[2] + udt) - (fabsf(udir.x) * hb[0] + fabsf(udir.y) * hb[1] + fabsf(udir.z) * hb[2])
                      - (udirA.x * ca[0] + udirA.y * ca[1] + udirA.z * ca[2]) - (fabsf(udirA.x) * ha[0] + fabsf(udirA.y) * ha[1] + fabsf(udirA.z) * ha[2])
                      - 4.0f * x.eps;
              if (g > 0.f) dc = fmaxf(dc, g * g);
            }
            bool lfl = oleaf && ch.child < 0;
            int mine = lfl ? ch.tri : cbase + c;
            int theirs = lfl ? other.tri : (splitA ? nb : na);
            int xa = splitA ? mine : theirs, xb = splitA ? theirs : mine;
            if (c == 0) { d0 = dc; i0a = xa; i0b = xb; lf0 = lfl; } else { d1 = dc; i1a = xa; i1b = xb; lf1 = lfl; }
          }
          n_box += 2;
          bool swap = d1 < d0;
          float dn = swap ? d1 : d0, df = swap ? d0 : d1;
          pn = dn <= best * DIST_BAND + 1.0e-12f; pf = df <= best * DIST_BAND + 1.0e-12f;
          ln = swap ? lf1 : lf0; lf = swap ? lf0 : lf1;
          cn = pack_dist_entry(dn, swap ? i1a : i0a, swap ? i1b : i0b);
          cf = pack_dist_entry(df, swap ? i0a : i1a, swap ? i0b : i1b);
        }
      }
      // far children first (deeper in the stack), near children on top
      unsigned sF = __ballot_sync(FULL, pf && !lf), sN = __ballot_sync(FULL, pn && !ln);
      unsigned qF = __ballot_sync(FULL, pf && lf), qN = __ballot_sync(FULL, pn && ln);
      if (pf && !lf) stack[sp + __popc(sF & lt_mask)] = cf;
      if (pn && !ln) stack[sp + __popc(sF) + __popc(sN & lt_mask)] = cn;
      sp += __popc(sF) + __popc(sN);
      if (pf && lf) leafq[lq + __popc(qF & lt_mask)] = cf;
      if (pn && ln) leafq[lq + __popc(qF) + __popc(qN & lt_mask)] = cn;
      lq += __popc(qF) + __popc(qN);
      __syncwarp();
    }
    __syncwarp();
    if (!hit && nnear > 0 && p.near_dir) near_emit(p, nearq, neard, nearf, nnear, best, best_fp, near_overflow, w);
    if (lane == 0) {
      if (hit) {
        p.dist[w] = -1.f;
        if (p.pts) for (int k = 0; k < 6; k++) p.pts[(size_t)w * 6 + k] = 0.0;
      } else {
        p.dist[w] = sqrtf(best);
        if (p.pts) {
          // bpa/bpb are in B's frame: p1 -> A's frame, p2 stays in B's frame
          d3 pa_local = xf_apply_inv(XA, xf_apply(XB, bpa));
          double *o = p.pts + (size_t)w * 6;
          o[0] = pa_local.x; o[1] = pa_local.y; o[2] = pa_local.z;
          o[3] = bpb.x; o[4] = bpb.y; o[5] = bpb.z;
        }
      }
    }
    __syncwarp();
  }
  if (p.stats) {
    unsigned long long v;
    v = n_box; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_DBOX], v);
    v = n_tri; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_DTRI], v);
#ifdef DIST_DEBUG_STATS
    v = dbg0; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_BOX], v);
    v = dbg1; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_TRI], v);
    v = dbg2; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_AMBIG], v);
    v = dbg3; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_HITFIX], v);
#endif
  }
}

// ============================================================================================
// fp64 refinement of tiny body-body distances: one warp per query that remembered near-minimum
// triangle pairs; lanes evaluate the pairs exactly (triangleTriangleDistanceSq arithmetic in fp64),
// the warp keeps the minimum.  Only queries whose distance is below ~1% of the triangles' extent get here.
// ============================================================================================
__global__ void __launch_bounds__(128) dist_refine_kernel(DistParams p) {
  const int lane = threadIdx.x & 31;
  int nseg = p.near_count[0];
  if (nseg > p.near_cap_dir) nseg = p.near_cap_dir;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nseg; s += (gridDim.x * blockDim.x) >> 5) {
    int4 d = p.near_dir[s];
    int w = d.x, ofs = d.y, cnt = d.z;
    if (cnt <= 0 || p.dist[w] < 0.f) continue;
    int pose = w / p.nq, qi = w - pose * p.nq;
    int2 ab = p.queries[qi];
    const MeshDev &ma = p.meshes[p.body_mesh[ab.x]];
    const MeshDev &mb = p.meshes[p.body_mesh[ab.y]];
    const double *XA = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.x);
    const double *XB = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.y);
    double bd = 1.0e300;
    d3 ba = mk<double>(0, 0, 0), bb = ba;
    for (int i = lane; i < cnt; i += 32) {
      int2 tt = p.near_ent[ofs + i];
      d3 a1 = tri_vertex(ma.tris, tt.x, 0), a2 = tri_vertex(ma.tris, tt.x, 1), a3 = tri_vertex(ma.tris, tt.x, 2);
      d3 b1 = tri_vertex(mb.tris, tt.y, 0), b2 = tri_vertex(mb.tris, tt.y, 1), b3 = tri_vertex(mb.tris, tt.y, 2);
      d3 P1 = xf_apply_inv(XB, xf_apply(XA, a1));
      d3 E1 = xf_rot_t(XB, xf_rot(XA, a2 - a1));
      d3 E2 = xf_rot_t(XB, xf_rot(XA, a3 - a2));
      d3 xa, xb;
      double dd = tri_tri_dist2<double>(P1, P1 + E1, P1 + E1 + E2, b1, b2, b3, xa, xb);
      if (dd < bd) { bd = dd; ba = xa; bb = xb; }
    }
    double m = bd;
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(FULL, m, o));
    unsigned who = __ballot_sync(FULL, bd == m);
    // the list holds every pair within tolerance of the fp32 best (flag 1): its fp64 minimum IS the answer; if the list
    // overflowed (flag 0) the fp32 answer stands unless a refined pair beats it
    const float cur = p.dist[w];
    // (flag 0: the list lost entries, so its minimum only counts if it is the fp32 winner itself -- equal within fp32
    // rounding -- or better)
    const bool replace = d.w != 0 || sqrt(m) <= (double)cur * (1.0 + 1.0e-6);
    if (replace && lane == __ffs(who) - 1 && m < 1.0e299) {
      p.dist[w] = (float)sqrt(m);
      if (p.pts) {
        d3 pa_local = xf_apply_inv(XA, xf_apply(XB, ba));
        double *o = p.pts + (size_t)w * 6;
        o[0] = pa_local.x; o[1] = pa_local.y; o[2] = pa_local.z;
        o[3] = bb.x; o[4] = bb.y; o[5] = bb.z;
      }
    }
  }
}

// ============================================================================================
// K5 + K8: CONTACT_ENERGY.  One thread per (legal pose, contact): every lane of a warp is busy with its
// own closest-point query, neighbouring lanes handle contacts of the same hand so their traversals
// stay close together in the tree.  A block owns `ppb` poses at a time (ppb * ncontacts <= blockDim),
// stores the per-contact terms in shared memory and sums them in a fixed order (deterministic).
// The top levels of the object's tree (a breadth-first prefix of the node array) are staged into shared
// memory once per block with one TMA bulk copy (cp.async.bulk + mbarrier).
//
// PRESET contacts come from the robot's virtual-contact table (reference robot.cpp:435-475).
// LIVE contacts are rebuilt from the body-body closest points exactly as the reference does, including
// its frame quirks: bodyToBodyDistance returns p1 = T_link^-1 * mP1,
// p2 = T_obj^-1 * mP2 with mP1/mP2 already local (graspitCollision.cpp:473-475); World::findVirtualContact
// (world.cpp:1641-1651) then forms n = normalize(T_link p1 - T_obj p2) = normalize(mP1 - mP2), the contact
// location is T_link p1 = mP1 read as a world point, and the virtual contact's normal is -n
// (virtualContact.cpp:92).
// ContactEnergy::energy (contactEnergy.cpp:9-55): mean over contacts of |v| + 50 (1 - cn . v/|v|).
// ============================================================================================
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// per-warp shared memory of the cooperative energy kernel
#ifndef EQ_STACK_PHYS_N
#define EQ_STACK_PHYS_N 512
#endif
#ifndef EQ_LEAF_TRIGGER
#define EQ_LEAF_TRIGGER 32      // queued leaves that start a batch of point-triangle tests
#endif
constexpr int EQ_STACK_MIN = EQ_STACK_PHYS_N, EQ_LEAF_CAP = 96;     // stack entries per warp: at least this, more for deep trees
constexpr unsigned long long EQ_NONE = 0x7f7fffffffffffffull;   // (FLT_MAX bits << 32) | no triangle
// per warp: leaf queue, best and runner-up keys, and the 32 queries in fp32 form (EnergyQueries below)
constexpr int ENERGY_WARP_FIXED = EQ_LEAF_CAP * 8 + 32 * 8 + 32 * 8 + 32 * 16 + 32 * 16 + 32 * 4;
// 32 depth-first traversals keep about one open sibling per level each
__host__ __device__ inline int energy_stack_entries(int depth) { (void)depth; return EQ_STACK_MIN; }   // deeper trees measured slower with larger stacks (occupancy)

__device__ __forceinline__ unsigned long long eq_entry(float lb, int q, int idx) {
  return ((unsigned long long)__float_as_uint(lb) << 32) | ((unsigned long long)(unsigned)q << 24) | (unsigned long long)(unsigned)idx;
}

// exact test of one triangle for one query -> key = squared distance | 8-bit hash of the closest point (1/64 mm grid) |
// triangle.  Triangles that share the closest vertex return bit-identical points: same hash, one solution, nothing to arbitrate.
__device__ __forceinline__ unsigned long long point_tri_key(const MeshDev &mo, d3 q, int tri) {
  d3 a = tri_vertex(mo.tris, tri, 0) - q, b = tri_vertex(mo.tris, tri, 1) - q, cc = tri_vertex(mo.tris, tri, 2) - q;
  f3 v = closest_pt_triangle<float>(tof(a), tof(b), tof(cc), mk<float>(0.f, 0.f, 0.f));
  float d2 = dot(v, v);
  const unsigned h = ((unsigned)__float2int_rd(v.x * 64.f) * 73856093u ^ (unsigned)__float2int_rd(v.y * 64.f) * 19349663u ^
                      (unsigned)__float2int_rd(v.z * 64.f) * 83492791u) & 0xffu;
  return ((unsigned long long)__float_as_uint(d2) << 32) | ((unsigned long long)h << 24) | (unsigned)tri;
}

// The 32 queries of a warp as the traversal reads them (shared memory).  A query point q (fp64, object frame) is held as
// the fp32 pair qh + ql (qh = fl(q), ql = fl(q - qh)): a triangle vertex relative to it is (v - qh) - ql in two fp32
// subtractions -- exact where it matters (v within a factor 2 of qh: Sterbenz), within half an ulp of the fp64
// difference rounded to fp32 elsewhere -- instead of 9 fp32->fp64 conversions, 9 fp64 subtractions and 9 conversions back
// on the quarter-rate pipe.  For the box tests the query also sits on the mesh's quantisation grid, rounded DOWN and
// biased by 2^23 (glo); see eq_gap2.
struct EnergyQueries {
  float4 *hx;        // [32] qh.x qh.y qh.z ql.x
  float4 *gl;        // [32] glo.x glo.y glo.z ql.y
  float *lz;         // [32] ql.z
};
#ifndef EQ_SPLIT_SUB
#define EQ_SPLIT_SUB 1
#endif
#ifndef EQ_GRID_GAP
#define EQ_GRID_GAP 1
#endif
__device__ __forceinline__ unsigned long long point_tri_key_s(const MeshDev &mo, f3 qh, f3 ql, int tri) {
#if EQ_SPLIT_SUB
  const float4 v0 = __ldg(&mo.tris[tri].v[0]), v1 = __ldg(&mo.tris[tri].v[1]), v2 = __ldg(&mo.tris[tri].v[2]);
  const f3 a = mk<float>((v0.x - qh.x) - ql.x, (v0.y - qh.y) - ql.y, (v0.z - qh.z) - ql.z);
  const f3 b = mk<float>((v1.x - qh.x) - ql.x, (v1.y - qh.y) - ql.y, (v1.z - qh.z) - ql.z);
  const f3 c = mk<float>((v2.x - qh.x) - ql.x, (v2.y - qh.y) - ql.y, (v2.z - qh.z) - ql.z);
  f3 v = closest_pt_triangle<float>(a, b, c, mk<float>(0.f, 0.f, 0.f));
  float d2 = dot(v, v);
  const unsigned h = ((unsigned)__float2int_rd(v.x * 64.f) * 73856093u ^ (unsigned)__float2int_rd(v.y * 64.f) * 19349663u ^
                      (unsigned)__float2int_rd(v.z * 64.f) * 83492791u) & 0xffu;
  return ((unsigned long long)__float_as_uint(d2) << 32) | ((unsigned long long)h << 24) | (unsigned)tri;
#else
  return point_tri_key(mo, mk<double>((double)qh.x + (double)ql.x, (double)qh.y + (double)ql.y, (double)qh.z + (double)ql.z), tri);
#endif
}
// Squared distance from a query to a stored box, straight from the 16-bit corners: 0x4B000000 | corner is the float
// 2^23 + corner (one byte permute, no integer-to-float conversion), the query's grid coordinate carries the same bias, so
// a gap is ONE exact subtraction of integers; only the three clamped gaps are scaled to millimetres.  The query is rounded
// down for the low faces and (glo + 2) up for the high faces, so the result never exceeds the true distance to the stored
// box (it falls short of it by at most two grid steps, a few micrometres): a lower bound, which is all the pruning needs.
__device__ __forceinline__ float eq_gap2(unsigned wx, unsigned wy, unsigned wz, float4 glo, const float *qs) {
  const float L0 = __uint_as_float(__byte_perm(wx, 0x4B000000u, 0x7610)), L1 = __uint_as_float(__byte_perm(wx, 0x4B000000u, 0x7632));
  const float L2 = __uint_as_float(__byte_perm(wy, 0x4B000000u, 0x7610)), H0 = __uint_as_float(__byte_perm(wy, 0x4B000000u, 0x7632));
  const float H1 = __uint_as_float(__byte_perm(wz, 0x4B000000u, 0x7610)), H2 = __uint_as_float(__byte_perm(wz, 0x4B000000u, 0x7632));
  const float g0 = fmaxf(fmaxf(L0 - (glo.x + 2.f), glo.x - H0), 0.f) * qs[0];
  const float g1 = fmaxf(fmaxf(L1 - (glo.y + 2.f), glo.y - H1), 0.f) * qs[1];
  const float g2 = fmaxf(fmaxf(L2 - (glo.z + 2.f), glo.z - H2), 0.f) * qs[2];
  return g0 * g0 + g1 * g1 + g2 * g2;
}
__device__ __forceinline__ float eq_box_dist2(const MeshDev &mo, uint4 w, float4 hx, float4 gl) {
#if EQ_GRID_GAP
  return eq_gap2(w.x, w.y, w.z, gl, mo.qs);
#else
  return point_nodeq_dist2(w, mo.q0, mo.qs, hx.x, hx.y, hx.z);
#endif
}
__device__ __forceinline__ void eq_store_query(const MeshDev &mo, const EnergyQueries &Q, int lane, d3 q) {
  const float hx = (float)q.x, hy = (float)q.y, hz = (float)q.z;
  const float lx = (float)(q.x - (double)hx), ly = (float)(q.y - (double)hy), lz = (float)(q.z - (double)hz);
  // grid coordinate, clamped to +-4e6 steps so that the biased value stays an exact integer (clamping moves the query
  // towards the boxes: still a lower bound), rounded down with a margin for the fp32 rounding of the division
  const float gx = fminf(fmaxf((hx - mo.q0[0]) / mo.qs[0], -4.0e6f), 4.0e6f);
  const float gy = fminf(fmaxf((hy - mo.q0[1]) / mo.qs[1], -4.0e6f), 4.0e6f);
  const float gz = fminf(fmaxf((hz - mo.q0[2]) / mo.qs[2], -4.0e6f), 4.0e6f);
  Q.hx[lane] = make_float4(hx, hy, hz, lx);
  Q.gl[lane] = make_float4(floorf(gx - 0.5f) + 8388608.f, floorf(gy - 0.5f) + 8388608.f, floorf(gz - 0.5f) + 8388608.f, ly);
  Q.lz[lane] = lz;
}

// contact c of a pose -> query point in the object's frame and contact normal in the object's frame (fp32)
__device__ __forceinline__ void energy_contact(const EnergyParams &p, int pose, int c, d3 &q, f3 &cnl) {
  const double *XO = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, p.object);
  d3 pw, cnw;
  if (p.live) {
    const double *o = p.pts + ((size_t)pose * p.nq + c) * 6;
    d3 mP1 = mk<double>(o[0], o[1], o[2]), mP2 = mk<double>(o[3], o[4], o[5]);
    d3 n = mP1 - mP2;
    double nn = sqrt(dot(n, n));
    if (nn > 0.0) n = n * (1.0 / nn);
    pw = mP1;
    cnw = mk<double>(-n.x, -n.y, -n.z);
  } else {
    const VcDev &vc = p.vcs[c];
    const double *XL = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, vc.body);
    pw = xf_apply(XL, mk<double>(vc.loc[0], vc.loc[1], vc.loc[2]));
    cnw = xf_rot(XL, mk<double>(vc.normal[0], vc.normal[1], vc.normal[2]));
  }
  q = xf_apply_inv(XO, pw);
  cnl = tof(xf_rot_t(XO, cnw));
}

// the energy term of one contact from the traversal's outcome: winner key (and the runner-up inside the near-tie band, if
// any).  The winning triangle is redone; in fp64 when the distance is tiny against its extent or a runner-up contests it.
__device__ __forceinline__ float energy_term_from_keys(const MeshDev &mo, d3 q, f3 cnl, unsigned long long key, unsigned long long key2) {
  if (key == EQ_NONE) return 0.f;
  const int tri = (int)(key & 0xffffffu);
  d3 a = tri_vertex(mo.tris, tri, 0) - q, b = tri_vertex(mo.tris, tri, 1) - q, cc = tri_vertex(mo.tris, tri, 2) - q;
  f3 af = tof(a), bf = tof(b), cf3 = tof(cc);
  f3 v = closest_pt_triangle<float>(af, bf, cf3, mk<float>(0.f, 0.f, 0.f));
  float d2 = dot(v, v);
  float sc = fmaxf(fmaxf(fmaxf(fabsf(af.x), fabsf(af.y)), fmaxf(fabsf(af.z), fabsf(bf.x))),
                   fmaxf(fmaxf(fabsf(bf.y), fabsf(bf.z)), fmaxf(fabsf(cf3.x), fmaxf(fabsf(cf3.y), fabsf(cf3.z)))));
  const bool contested = key2 != EQ_NONE && ((key2 ^ key) & 0xff000000ull) != 0ull &&
                         __uint_as_float((unsigned)(key2 >> 32)) <= d2 * NEAR_BAND + 1.0e-12f;
  if (d2 < 1.0e-4f * sc * sc || contested) {
    d3 v64 = closest_pt_triangle_d(a, b, cc);
    if (contested) {
      const int tri2 = (int)(key2 & 0xffffffu);
      d3 a2 = tri_vertex(mo.tris, tri2, 0) - q, b2 = tri_vertex(mo.tris, tri2, 1) - q, c2 = tri_vertex(mo.tris, tri2, 2) - q;
      d3 w64 = closest_pt_triangle_d(a2, b2, c2);
      if (dot(w64, w64) < dot(v64, v64)) v64 = w64;
    }
    v = tof(v64);
    d2 = (float)dot(v64, v64);
  }
  const float dist = sqrtf(d2);
  const float inv = d2 > 0.f ? 1.0f / dist : 0.f;
  return dist + (1.0f - (cnl.x * v.x + cnl.y * v.y + cnl.z * v.z) * inv) * 50.0f;
}

// exact test of one leaf for query qi of the warp: best (distance, triangle) by 64-bit atomicMin; the best loser with a
// DIFFERENT closest point inside the near-tie band is remembered for the fp64 arbitration of the final stage (fp32 cannot
// order two candidates whose distances agree to 1e-4, and the energy depends on which closest point wins)
__device__ __forceinline__ void eq_leaf_test(const MeshDev &mo, const EnergyQueries &Q, unsigned long long *bestkey, unsigned long long *secondkey,
                                             int qi, int tri) {
  const float4 hx = Q.hx[qi], gl = Q.gl[qi];
  const unsigned long long key = point_tri_key_s(mo, mk<float>(hx.x, hx.y, hx.z), mk<float>(hx.w, gl.w, Q.lz[qi]), tri);
  const unsigned long long old = atomicMin(&bestkey[qi], key);
  const unsigned long long win = old < key ? old : key, lose = old < key ? key : old;
  if (lose != EQ_NONE && ((win ^ lose) & 0xff000000ull) != 0ull &&
      __uint_as_float((unsigned)(lose >> 32)) <= __uint_as_float((unsigned)(win >> 32)) * NEAR_BAND + 1.0e-12f)
    atomicMin(&secondkey[qi], lose);
}

// The shared stack is full: one lane finishes the subtree of the top entry on a private stack (depth-first, nearer child
// first, same pruning, same leaf test).  Rare -- 32 depth-first traversals keep about 32 x depth entries open and the stack
// is sized for that -- but it makes the traversal exact whatever the tree depth and however the queries interleave.
// (returns box tests | triangle tests << 16 for the statistics: counters passed by reference would live in memory)
__device__ __noinline__ unsigned energy_private_dfs(const MeshDev &mo, const NodeQ *top, int ntop, unsigned long long e, const EnergyQueries &Q,
                                                    unsigned long long *bestkey, unsigned long long *secondkey) {
  unsigned nbx = 0, ntr = 0;
  const int qi = (int)((e >> 24) & 0xffu);
  int stk[64]; float stl[64];
  int n = 0;
  stk[n] = (int)(e & 0xffffffu); stl[n] = __uint_as_float((unsigned)(e >> 32)); n++;
  const float4 qhx = Q.hx[qi], qgl = Q.gl[qi];
  while (n > 0) {
    n--;
    const int pair = stk[n];
    float best = __uint_as_float((unsigned)(bestkey[qi] >> 32));
    if (stl[n] > best * NEAR_BAND + 1.0e-12f) continue;
    uint4 w0, w1;
    fetch_pair_raw(mo, top, ntop, pair, w0, w1);
    float d0 = eq_box_dist2(mo, w0, qhx, qgl);
    float d1 = eq_box_dist2(mo, w1, qhx, qgl);
    nbx += 2;
    const bool swap = d1 < d0;
    const float dn = swap ? d1 : d0, df = swap ? d0 : d1;
    const int lkn = (int)(swap ? w1.w : w0.w), lkf = (int)(swap ? w0.w : w1.w);
    // far child first (deeper in the stack), leaves tested at once
    if (df <= best * NEAR_BAND + 1.0e-12f) {
      if (lkf < 0) { eq_leaf_test(mo, Q, bestkey, secondkey, qi, ~lkf); ntr++; }
      else if (n < 64) { stk[n] = lkf; stl[n] = df; n++; }
    }
    best = __uint_as_float((unsigned)(bestkey[qi] >> 32));
    if (dn <= best * NEAR_BAND + 1.0e-12f) {
      if (lkn < 0) { eq_leaf_test(mo, Q, bestkey, secondkey, qi, ~lkn); ntr++; }
      else if (n < 64) { stk[n] = lkn; stl[n] = dn; n++; }
    }
  }
  return (nbx & 0xffffu) | (ntr << 16);
}

// One warp = 32 closest-point queries (contacts of consecutive legal poses) traversed TOGETHER through one
// shared work stack: every lane pops one (query, child-pair, lower bound) entry, loads the 64-byte pair of
// sibling boxes, and pushes the children that can still beat the query's best distance (far first, near on
// top, so each query advances depth-first, nearest first); finished queries simply stop producing entries
// and their lanes pick up other queries' work.  Leaves are queued and tested 32 triangles at a time; the
// per-query best (distance, triangle) is a 64-bit atomicMin in shared memory.
#ifndef ENERGY_MIN_BLOCKS
#define ENERGY_MIN_BLOCKS 4      // 64 registers: 4 blocks (32 warps) per SM; measured best against 3 x 80 registers
#endif
__global__ void __launch_bounds__(256, ENERGY_MIN_BLOCKS) energy_kernel(EnergyParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NodeQ *top = reinterpret_cast<NodeQ *>(smem_raw);
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int stack_cap = p.stack_entries;
  unsigned char *wbase = smem_raw + (size_t)p.stage_nodes * sizeof(NodeQ) + (size_t)warp * (ENERGY_WARP_FIXED + 8 * stack_cap);
  unsigned long long *stack = reinterpret_cast<unsigned long long *>(wbase);
  unsigned long long *leafq = stack + stack_cap;
  unsigned long long *bestkey = leafq + EQ_LEAF_CAP;                 // [32] (d2 bits << 32) | point hash << 24 | triangle
  unsigned long long *secondkey = bestkey + 32;                       // [32] best loser within the near-tie band of the winner
  EnergyQueries Q;
  Q.hx = reinterpret_cast<float4 *>(secondkey + 32);
  Q.gl = Q.hx + 32;
  Q.lz = reinterpret_cast<float *>(Q.gl + 32);
  const int nc = p.live ? p.nq : p.nvc;
  const MeshDev mo = p.meshes[p.body_mesh[p.object]];
  const int ntop = p.stage_nodes < mo.nnodes ? p.stage_nodes : mo.nnodes;

  if (ntop > 0) {
    // TMA bulk copy of the tree's top levels (breadth-first prefix) into shared memory
    const unsigned bytes = (unsigned)ntop * (unsigned)sizeof(NodeQ);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(top)), "l"(mo.nodes), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(&bar)) : "memory");
  }

  const int nlegal = p.legal_count ? *p.legal_count : p.npose;
  const int nquery = nlegal * nc;
  const int ngroups = (nquery + 31) / 32;
  unsigned nbx = 0, ntr = 0;
  const Node root = fetch_node(mo, top, ntop, 0);

  while (true) {
    // groups of 32 queries are handed out dynamically, so the grid need not match the resident block count
    int g = 0;
    if (lane == 0) g = atomicAdd(p.work_counter, 1);
    g = __shfl_sync(FULL, g, 0);
    if (g >= ngroups) break;
    // ---- set up this lane's query
    const int qid = g * 32 + lane;
    bool valid = qid < nquery;
    int pose = -1, c = 0;
    f3 cnl = mk<float>(0.f, 0.f, 0.f);
    unsigned long long seedkey = EQ_NONE;
    if (valid) {
      int li = qid / nc;
      c = qid - li * nc;
      pose = p.legal_list ? p.legal_list[li] : li;
      if (p.legal && !p.legal[pose]) valid = false;
    }
    d3 q = mk<double>(0, 0, 0);
    if (valid) {
      energy_contact(p, pose, c, q, cnl);
      eq_store_query(mo, Q, lane, q);
      if (mo.seed) {
        // start from a real candidate: the triangle closest to the centre of the query's grid cell (clamped into the
        // grid for far queries), tested exactly like any leaf.  Its key is what the traversal would have produced had it
        // met that triangle first, so the result is unchanged; what changes is how much of the tree survives the bound.
        const int ix = min(max(__float2int_rd(((float)q.x - mo.g0[0]) * mo.ginv[0]), 0), mo.gres[0] - 1);
        const int iy = min(max(__float2int_rd(((float)q.y - mo.g0[1]) * mo.ginv[1]), 0), mo.gres[1] - 1);
        const int iz = min(max(__float2int_rd(((float)q.z - mo.g0[2]) * mo.ginv[2]), 0), mo.gres[2] - 1);
        const int tri = __ldg(&mo.seed[(iz * mo.gres[1] + iy) * mo.gres[0] + ix]);
        if (tri >= 0) {
          const float4 hx = Q.hx[lane], gl = Q.gl[lane];
          seedkey = point_tri_key_s(mo, mk<float>(hx.x, hx.y, hx.z), mk<float>(hx.w, gl.w, Q.lz[lane]), tri);
          ntr++;
        }
      }
    }
    bestkey[lane] = seedkey;
    secondkey[lane] = EQ_NONE;
    int sp = 0, lq = 0;
    {
      // seed: the root's child pair (or the root leaf itself for one-triangle meshes)
      bool inner = valid && root.child >= 0, leafroot = valid && root.child < 0 && root.tri >= 0;
      unsigned mi = __ballot_sync(FULL, inner), ml = __ballot_sync(FULL, leafroot);
      if (inner) stack[__popc(mi & lt_mask)] = eq_entry(0.f, lane, root.child);
      if (leafroot) leafq[__popc(ml & lt_mask)] = eq_entry(0.f, lane, root.tri);
      sp = __popc(mi); lq = __popc(ml);
    }
    __syncwarp();

    while (true) {
      if (lq >= EQ_LEAF_TRIGGER || (sp == 0 && lq > 0)) {
        // ---- up to 32 point-triangle tests
        int n = lq < 32 ? lq : 32;
        if (lane < n) {
          unsigned long long e = leafq[lq - n + lane];
          int qi = (int)((e >> 24) & 0xffu), tri = (int)(e & 0xffffffu);
          float lb = __uint_as_float((unsigned)(e >> 32));
          float best = __uint_as_float((unsigned)(bestkey[qi] >> 32));
          if (lb <= best * NEAR_BAND + 1.0e-12f) {
            ntr++;
            eq_leaf_test(mo, Q, bestkey, secondkey, qi, tri);
          }
        }
        lq -= n;
        __syncwarp();
        continue;
      }
      if (sp == 0) break;
      // k entries popped push at most 2k.  Above the soft limit only one entry is popped per round (the depth-first order
      // then shrinks the stack again); should it still reach the physical size, one lane finishes the top entry's subtree
      // on a private stack, so the shared one is never outgrown
      const int room = stack_cap - 64 - sp;
      if (sp >= stack_cap - 2) {
        if (lane == 0) { const unsigned c = energy_private_dfs(mo, top, ntop, stack[sp - 1], Q, bestkey, secondkey); nbx += c & 0xffffu; ntr += c >> 16; }
        sp -= 1;
        __syncwarp();
        continue;
      }
      int k = sp < 32 ? sp : 32;
      if (k > room) k = room > 1 ? room : 1;
      bool have = lane < k;
      unsigned long long e = have ? stack[sp - 1 - lane] : 0ull;
      sp -= k;
      __syncwarp();
      bool pn = false, pf = false, ln = false, lf = false;
      unsigned long long cn = 0ull, cf = 0ull;
      if (have) {
        int qi = (int)((e >> 24) & 0xffu), pair = (int)(e & 0xffffffu);
        float lb = __uint_as_float((unsigned)(e >> 32));
        float best = __uint_as_float((unsigned)(bestkey[qi] >> 32));
        if (lb <= best * NEAR_BAND + 1.0e-12f) {
          uint4 w0, w1;
          fetch_pair_raw(mo, top, ntop, pair, w0, w1);
          const float4 qhx = Q.hx[qi], qgl = Q.gl[qi];
          float d0 = eq_box_dist2(mo, w0, qhx, qgl);
          float d1 = eq_box_dist2(mo, w1, qhx, qgl);
          nbx += 2;
          bool swap = d1 < d0;
          float dn = swap ? d1 : d0, df = swap ? d0 : d1;
          // link >= 0: child pair index; link < 0: leaf holding triangle ~link
          int lkn = (int)(swap ? w1.w : w0.w), lkf = (int)(swap ? w0.w : w1.w);
          int chn = lkn >= 0 ? lkn : -1, chf = lkf >= 0 ? lkf : -1;
          int trn = ~lkn, trf = ~lkf;
          pn = dn <= best * NEAR_BAND + 1.0e-12f; pf = df <= best * NEAR_BAND + 1.0e-12f;
          ln = chn < 0; lf = chf < 0;
          cn = eq_entry(dn, qi, ln ? trn : chn);
          cf = eq_entry(df, qi, lf ? trf : chf);
        }
      }
      unsigned sF = __ballot_sync(FULL, pf && !lf), sN = __ballot_sync(FULL, pn && !ln);
      unsigned qF = __ballot_sync(FULL, pf && lf), qN = __ballot_sync(FULL, pn && ln);
      if (pf && !lf) stack[sp + __popc(sF & lt_mask)] = cf;
      if (pn && !ln) stack[sp + __popc(sF) + __popc(sN & lt_mask)] = cn;
      sp += __popc(sF) + __popc(sN);
      if (pf && lf) leafq[lq + __popc(qF & lt_mask)] = cf;
      if (pn && ln) leafq[lq + __popc(qF) + __popc(qN & lt_mask)] = cn;
      lq += __popc(qF) + __popc(qN);
      __syncwarp();
    }
    __syncwarp();

    // ---- this lane's result: redo the winning triangle (fp64 when the distance is tiny against its extent)
    if (valid) {
      // qh + ql gives q back to 2^-48 of |q - qh| (~1e-20 mm): not worth six registers across the traversal
      const float4 hx = Q.hx[lane], gl = Q.gl[lane];
      const d3 qq = mk<double>((double)hx.x + (double)hx.w, (double)hx.y + (double)gl.w, (double)hx.z + (double)Q.lz[lane]);
      p.terms[qid] = energy_term_from_keys(mo, qq, cnl, bestkey[lane], secondkey[lane]);
    }
    __syncwarp();
  }
  if (p.stats) {
    unsigned long long v;
    v = nbx; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_PBOX], v);
    v = ntr; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_PTRI], v);
  }
}

// closest-point seed grid of a mesh: one thread per cell, the triangle closest to the cell's centre (per-lane traversal,
// run once per mesh)
__global__ void __launch_bounds__(128) seed_grid_kernel(MeshDev m, int *seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = m.gres[0] * m.gres[1] * m.gres[2];
  if (i >= n) return;
  const int ix = i % m.gres[0], iy = (i / m.gres[0]) % m.gres[1], iz = i / (m.gres[0] * m.gres[1]);
  d3 q = mk<double>((double)m.g0[0] + ((double)ix + 0.5) / (double)m.ginv[0], (double)m.g0[1] + ((double)iy + 0.5) / (double)m.ginv[1],
                    (double)m.g0[2] + ((double)iz + 0.5) / (double)m.ginv[2]);
  f3 v; int bt = -1; unsigned nb = 0, nt = 0;
  closest_point_query(m, nullptr, 0, q, v, bt, nb, nt);
  seed[i] = bt;
}
cudaError_t launch_seed_grid(const MeshDev &m, int *seed, cudaStream_t s) {
  const int n = m.gres[0] * m.gres[1] * m.gres[2];
  if (n <= 0) return cudaSuccess;
  seed_grid_kernel<<<(n + 127) / 128, 128, 0, s>>>(m, seed);
  return cudaGetLastError();
}

// fixed-order sum of the per-contact terms of every legal pose (ContactEnergy::energy's division by the
// contact count, contactEnergy.cpp:51)
__global__ void energy_reduce_kernel(EnergyParams p) {
  const int nc = p.live ? p.nq : p.nvc;
  const int nlegal = p.legal_count ? *p.legal_count : p.npose;
  for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < nlegal; li += gridDim.x * blockDim.x) {
    int pose = p.legal_list ? p.legal_list[li] : li;
    if (p.legal && !p.legal[pose]) { p.energy[pose] = 0.f; continue; }
    float total = 0.f;
    for (int k = 0; k < nc; k++) total += p.terms[(size_t)li * nc + k];
    p.energy[pose] = total / (float)nc;
  }
}

// ---------------------------------------------------------------------------------------------
// host-callable launchers (api.cu)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_fk(const FkParams &p, cudaStream_t s) {
  if (p.npose <= 0) return cudaSuccess;
  fk_kernel<<<(p.npose + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_recheck(const RecheckParams &p, cudaStream_t s) {
  recheck_kernel<<<64, 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_point(const PointParams &p, cudaStream_t s) {
  if (p.n <= 0) return cudaSuccess;
  point_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_dist(const DistParams &p, int blocks, cudaStream_t s) {
  size_t smem = (size_t)8 * DIST_WARP_SMEM;
  static size_t granted[64] = {};
  { cudaError_t e = ensure_dyn_smem((const void *)dist_kernel, smem, granted); if (e != cudaSuccess) return e; }
  dist_kernel<<<blocks, 256, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !p.near_dir) return e;
  dist_refine_kernel<<<blocks > 1 ? 148 * 8 : 8, 128, 0, s>>>(p);     // one warp per near-tie query, grid-stride
  return cudaGetLastError();
}

cudaError_t launch_energy(const EnergyParams &p, int blocks, cudaStream_t s) {
  if (p.npose <= 0) return cudaSuccess;
  size_t smem = energy_block_smem_bytes(p.stage_nodes, p.stack_entries);
  static size_t granted[64] = {};
  { cudaError_t e = ensure_dyn_smem((const void *)energy_kernel, smem, granted); if (e != cudaSuccess) return e; }
  energy_kernel<<<blocks, 256, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  energy_reduce_kernel<<<(p.npose + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

size_t energy_block_smem_bytes(int stage_nodes, int stack_entries) {
  return (size_t)stage_nodes * sizeof(NodeQ) + (size_t)8 * (ENERGY_WARP_FIXED + 8 * (size_t)stack_entries);
}
int energy_stack_entries_for(int depth) { return energy_stack_entries(depth); }

} // namespace b2c
