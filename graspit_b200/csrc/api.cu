// C ABI of the engine (include/b2c.h): context, geometry registry, pair activation, robots,
// single-pose CollisionInterface queries and the batched FK -> legal -> energy evaluation.
// Host-side bookkeeping mirrors GraspitCollision (src/Collision/Graspit/graspitCollision.cpp):
// body <-> model map, disabled-pair set, active-pair enumeration (:55-100).
#include "../../include/b2c.h"
#include "b2c_internal.h"
#include "contacts_host.h"
#include "obb_order.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace b2c;

namespace {

constexpr int B2C_MAX_CHUNKS = 8;

struct MeshHost {
  bool alive = false;
  int ntri = 0;
  std::vector<float> tri9;        // kept for the host-side OBB hierarchy / region / contact post-processing
  std::vector<NodeQ> nodes;
  BvhQuant quant;
  NodeQ *d_nodes = nullptr;
  Tri *d_tris = nullptr;
  int depth = 0;
  float extent = 0.f;
  ObbTree *obb = nullptr;         // lazily built reference-style hierarchy (getBoundingVolumes)
  OrderNode *d_order = nullptr;   // the same tree as the visit-order kernel reads it (contact_order.cu), uploaded on first use
  int *d_tri_pos = nullptr;
  int *d_seed = nullptr;          // lazily built closest-point seed grid (meshes used as the energy object)
  float g0[3] = {0, 0, 0}, ginv[3] = {0, 0, 0};
  int gres[3] = {0, 0, 0};
  float root_c[3] = {0, 0, 0}, root_h[3] = {0, 0, 0};
};

struct BodyHost {
  bool alive = false;
  int mesh = -1;
  bool active = true;
  int thread = 0;                 // CollisionModel::getThreadId (collisionModel.h): 0 = master thread
  Qt tf;
};

struct RobotHost {
  b2c_robot_desc d;
  std::vector<double> dof_min, dof_max, eg, eg_origin, eg_norm;
  std::vector<b2c_chain_desc> chains;
  std::vector<b2c_joint_desc> joints;
  std::vector<int> link_body, link_last_joint;
  std::vector<b2c_vcontact_desc> vcs;
};

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Plan {
  bool valid = false;
  int robot = -1, object = -1;
  std::vector<int> robots;          // robot ids in FK order (root first)
  std::vector<int> slots;           // robot body ids, slot order
  std::vector<int> ebodies;         // eval bodies: slots, then static bodies
  std::vector<int2> pairs;          // eval-body indices
  std::vector<int2> pair_ids;       // body ids
  int object_e = -1;
  int ndof_total = 0;
  int nvc = 0;
  int nhand = 0;                    // finger links + palm of the root robot (LIVE contacts / link_dist); the mount is not one
  // device copies
  DevBuf<FkRobot> d_robots; DevBuf<FkChain> d_chains; DevBuf<FkJoint> d_joints; DevBuf<FkLink> d_links;
  DevBuf<double> d_eg; DevBuf<int> d_body_mesh; DevBuf<int2> d_pairs; DevBuf<VcDev> d_vcs; DevBuf<int2> d_queries;
  DevBuf<Xf> d_static;
  std::vector<Xf> static_cache;     // what d_static holds
  FkProgram prog;
  void release() {
    d_robots.release(); d_chains.release(); d_joints.release(); d_links.release(); d_eg.release();
    d_body_mesh.release(); d_pairs.release(); d_vcs.release(); d_queries.release(); d_static.release();
    valid = false;
  }
};

} // namespace

struct AnnealHost;      // anneal_api.inc
struct ExchangeHost;    // anneal_api.inc
namespace { struct AgScratch; }       // autograsp_api.inc

struct b2c_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  int sm_count = 148;
  std::vector<MeshHost> meshes;
  std::vector<BodyHost> bodies;
  std::set<std::pair<int, int>> disabled;
  int query_thread = 0;            // CollisionInterface::getThreadId of the caller (b2c_set_thread)
  std::vector<RobotHost> robots;
  bool meshes_dirty = true;
  DevBuf<MeshDev> d_meshes;
  std::map<std::pair<int, int>, Plan *> plans;
  // scratch
  DevBuf<float> d_states; DevBuf<Xf> d_fk; DevBuf<double> d_fkqt; DevBuf<uint8_t> d_legal; DevBuf<float> d_energy;
  DevBuf<unsigned long long> d_mask; DevBuf<Ambig> d_ambig; DevBuf<int> d_counters; DevBuf<unsigned long long> d_stats;
  DevBuf<float> d_dist; DevBuf<double> d_pts; DevBuf<double> d_misc; DevBuf<int> d_imisc; DevBuf<int> d_legal_list; DevBuf<int2> d_items; DevBuf<ContactRec> d_contacts; DevBuf<unsigned char> d_flags; DevBuf<float> d_terms; DevBuf<int4> d_near_dir; DevBuf<int2> d_near_ent;
  ContactRec *h_contacts[2] = {nullptr, nullptr}; // pinned staging of contact candidates, double-buffered (b2c_contacts_batch)
  DevBuf<unsigned long long> d_okeys; DevBuf<int> d_odepth; DevBuf<int> d_operm; DevBuf<unsigned long long> d_order_tab;
  DevBuf<int> d_nsurv, d_soff, d_dtotals, d_leader, d_leaders;                // device duplicate removal of a batched contacts call
  int *h_nsurv = nullptr; size_t h_nsurv_cap = 0; int *h_dtotals = nullptr;   // pinned
  cudaStream_t copy_stream2 = nullptr; cudaEvent_t ev_surv = nullptr;
  int2 *h_hits = nullptr; size_t h_hits_cap = 0;        // pinned: flagged (triangle, query) pairs of a region batch
  DevBuf<int2> d_items2;
  int *h_operm = nullptr; size_t h_operm_cap = 0;       // pinned: the candidates' order inside their groups
  cudaStream_t copy_stream = nullptr;              // D2H copies of a batched contacts call, beside the compute stream
  cudaEvent_t ev_keys = nullptr;
  double contacts_per_pose_seen = 0.0;             // candidate records per pose of the densest batched contacts call so far
  cudaEvent_t ev_chunk[B2C_MAX_CHUNKS] = {};       // one per chunk of a batched contacts call: "this chunk's copies have landed"
  DevBuf<ContactRec> d_contacts2;   // the same records grouped by query
  DevBuf<int> d_coffsets;           // per query: first record, fill cursor
  DevBuf<unsigned char> d_cubtemp;
  std::vector<int> h_soff;
  DevBuf<int> d_rfirst, d_rkeep, d_rpos, d_roff; DevBuf<double> d_rout;     // vertex lists of a region batch
  void *h_rout = nullptr; size_t h_rout_cap = 0;                           // pinned staging of the finished lists
  int *h_qoff = nullptr; size_t h_qoff_cap = 0;        // pinned: first record of every (pose, pair) query of a batched contacts call
  int *h_qsmall = nullptr; size_t h_qsmall_cap = 0;    // pinned: chunk starts + pairs in use (contact_digest_kernel)
  DevBuf<int> d_qsmall;
  size_t h_contacts_cap[2] = {0, 0};
  Xf *h_fk = nullptr;               // pinned copy of the link transforms (reference-order replay of contact candidates)
  size_t h_fk_cap = 0;
  int stage_nodes = 128;           // object-tree nodes staged in shared memory by the energy kernel
  int seed_res = 64;               // cells along the longest side of a closest-point seed grid (0 = no seeding)
  unsigned long long *last_stats = nullptr;    // pinned: the per-call copy must not block the host (a pageable D2H does)
  int last_counters[4] = {0, 0, 0, 0};
  int eval_warps_per_block = 8;
  int eval_blocks_per_sm = 0;      // 0 = derive from shared memory
  cudaStream_t own_stream = nullptr;
  // per-context annealers and autograsp scratch (contexts are independent: one per GPU, possibly one host thread each)
  std::vector<AnnealHost *> anneals;
  ExchangeHost *exchange = nullptr;       // best-row exchange with the other GPUs (b2c_exchange_*)
  AgScratch *autograsp = nullptr;
  // sticky overflow word on the device: bit 0 = legality recheck queue, bit 1 = distance recheck queue, bit 2 = broad-phase
  // work-item list.  Kernels that drop an entry set it (and answer conservatively: "collides"); synchronising calls report it.
  DevBuf<int> d_sticky;
  bool profiling = false;
  cudaEvent_t ev[16] = {};
  bool ev_used[8] = {};
};

namespace {

int fail(b2c_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(ctx, B2C_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

bool body_ok(const b2c_ctx *c, int b) { return b >= 0 && b < (int)c->bodies.size() && c->bodies[b].alive; }
bool mesh_ok(const b2c_ctx *c, int m) { return m >= 0 && m < (int)c->meshes.size() && c->meshes[m].alive; }

Qt make_qt(const double *q, const double *t) {
  Qt r; r.w = q[0]; r.x = q[1]; r.y = q[2]; r.z = q[3]; r.tx = t[0]; r.ty = t[1]; r.tz = t[2];
  q_normalize(r);
  return r;
}
Xf qt_to_xf(const Qt &q) { Xf x; q_to_xf(q, reinterpret_cast<double *>(&x)); return x; }

void invalidate_plans(b2c_ctx *c) {
  for (auto &kv : c->plans) { kv.second->release(); delete kv.second; }
  c->plans.clear();
}

int sync_meshes(b2c_ctx *ctx) {
  if (!ctx->meshes_dirty) return B2C_OK;
  std::vector<MeshDev> md(ctx->meshes.size());
  for (size_t i = 0; i < ctx->meshes.size(); i++) {
    const MeshHost &m = ctx->meshes[i];
    md[i].nodes = m.d_nodes; md[i].tris = m.d_tris;
    md[i].nnodes = (int)m.nodes.size(); md[i].ntri = m.ntri;
    md[i].extent = m.extent; md[i].depth = m.depth;
    for (int k = 0; k < 3; k++) { md[i].q0[k] = m.quant.q0[k]; md[i].qs[k] = m.quant.qs[k]; }
    md[i].seed = m.d_seed;
    for (int k = 0; k < 3; k++) { md[i].g0[k] = m.g0[k]; md[i].ginv[k] = m.ginv[k]; md[i].gres[k] = m.gres[k]; }
  }
  CU(ctx->d_meshes.reserve(md.size() + 1));
  if (!md.empty()) CU(cudaMemcpyAsync(ctx->d_meshes.p, md.data(), md.size() * sizeof(MeshDev), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->meshes_dirty = false;
  return B2C_OK;
}

// thread rule of getActivePairs (graspitCollision.cpp:64,77-78): a body takes part when it belongs to the
// querying thread or to the master thread (0); a non-master thread never tests two master bodies
bool thread_visible(const b2c_ctx *c, int a) { return c->bodies[a].thread == c->query_thread || c->bodies[a].thread == 0; }
bool thread_pair_ok(const b2c_ctx *c, int a, int b) {
  if (!thread_visible(c, a) || !thread_visible(c, b)) return false;
  return !(c->query_thread != 0 && c->bodies[a].thread == 0 && c->bodies[b].thread == 0);
}

bool pair_active(const b2c_ctx *c, int a, int b) {
  if (a == b) return false;
  if (!c->bodies[a].active || !c->bodies[b].active) return false;   // isActive() ignores threads (graspitCollision.cpp:158-198)
  return c->disabled.find({std::min(a, b), std::max(a, b)}) == c->disabled.end();
}

// GraspitCollision::getActivePairs (graspitCollision.cpp:55-100): both active, not disabled, at least
// one body in the interest list (NULL = all).  The reference orders pairs by model pointer; we order
// by body id, and callers compare as sets.
void active_pairs(const b2c_ctx *c, const std::set<int> *interest, std::vector<std::pair<int, int>> &out) {
  out.clear();
  int n = (int)c->bodies.size();
  for (int a = 0; a < n; a++) {
    if (!c->bodies[a].alive || !c->bodies[a].active) continue;
    bool ia = !interest || interest->count(a);
    for (int b = a + 1; b < n; b++) {
      if (!c->bodies[b].alive || !c->bodies[b].active) continue;
      if (!ia && !interest->count(b)) continue;
      if (!thread_pair_ok(c, a, b)) continue;
      if (c->disabled.count({a, b})) continue;
      out.push_back({a, b});
    }
  }
}

// Closest-point seed grid of the mesh an energy evaluation queries (built once per mesh, on the device): the mesh's
// bounding box grown by a quarter of its longest side, cubic cells, seed_res cells along the longest side.
int ensure_seed_grid(b2c_ctx *ctx, int mesh) {
  MeshHost &m = ctx->meshes[mesh];
  if (m.d_seed || ctx->seed_res <= 0 || m.ntri <= 0) return B2C_OK;
  const float longest = 2.f * std::max(m.root_h[0], std::max(m.root_h[1], m.root_h[2]));
  const float margin = 0.25f * longest;
  const float cell = (longest + 2.f * margin) / (float)ctx->seed_res;
  size_t cells = 1;
  for (int k = 0; k < 3; k++) {
    const float side = 2.f * m.root_h[k] + 2.f * margin;
    m.gres[k] = std::max(1, (int)std::ceil(side / cell));
    m.g0[k] = m.root_c[k] - 0.5f * (float)m.gres[k] * cell;
    m.ginv[k] = 1.f / cell;
    cells *= (size_t)m.gres[k];
  }
  CU(cudaMalloc(&m.d_seed, cells * sizeof(int)));
  MeshDev md;
  memset(&md, 0, sizeof md);
  md.nodes = m.d_nodes; md.tris = m.d_tris; md.nnodes = (int)m.nodes.size(); md.ntri = m.ntri; md.extent = m.extent; md.depth = m.depth;
  for (int k = 0; k < 3; k++) { md.q0[k] = m.quant.q0[k]; md.qs[k] = m.quant.qs[k]; md.g0[k] = m.g0[k]; md.ginv[k] = m.ginv[k]; md.gres[k] = m.gres[k]; }
  CU(launch_seed_grid(md, m.d_seed, ctx->stream)); ctx->launches++;
  ctx->meshes_dirty = true;
  return B2C_OK;
}

int ensure_scratch(b2c_ctx *ctx) {
  if (!ctx->d_sticky.p) {
    CU(ctx->d_sticky.reserve(4));
    CU(cudaMemsetAsync(ctx->d_sticky.p, 0, 4 * sizeof(int), ctx->stream));
  }
  CU(ctx->d_counters.reserve(16));
  CU(ctx->d_stats.reserve(8));
  CU(ctx->d_ambig.reserve(1 << 20));
  CU(ctx->d_near_dir.reserve(1 << 18));
  CU(ctx->d_near_ent.reserve(1 << 21));
  return B2C_OK;
}

// Run the collision traversal over explicit static bodies (single-pose CollisionInterface queries).
int run_static_collisions(b2c_ctx *ctx, const std::vector<std::pair<int, int>> &pairs, bool all_mode,
                          std::vector<unsigned long long> &mask, int &legal) {
  legal = 1;
  mask.assign((pairs.size() + 63) / 64 + 1, 0ull);
  if (pairs.empty()) return B2C_OK;
  int rc = sync_meshes(ctx); if (rc) return rc;
  rc = ensure_scratch(ctx); if (rc) return rc;
  std::map<int, int> eidx;
  std::vector<int> ebodies;
  std::vector<int2> epairs;
  for (auto &pr : pairs) {
    for (int b : {pr.first, pr.second})
      if (!eidx.count(b)) { eidx[b] = (int)ebodies.size(); ebodies.push_back(b); }
    epairs.push_back(make_int2(eidx[pr.first], eidx[pr.second]));
  }
  int nb = (int)ebodies.size(), np = (int)epairs.size();
  std::vector<Xf> sx(nb);
  std::vector<int> bm(nb);
  for (int i = 0; i < nb; i++) { sx[i] = qt_to_xf(ctx->bodies[ebodies[i]].tf); bm[i] = ctx->bodies[ebodies[i]].mesh; }
  int mw = (np + 63) / 64;
  size_t bytes_x = sizeof(Xf) * nb, bytes_p = sizeof(int2) * np, bytes_m = sizeof(int) * nb;
  CU(ctx->d_misc.reserve((bytes_x + 7) / 8 + 8));
  CU(ctx->d_imisc.reserve(bytes_p / 4 + bytes_m / 4 + 16));
  CU(ctx->d_mask.reserve(mw + 1));
  CU(ctx->d_legal.reserve(16));
  int2 *d_pairs = reinterpret_cast<int2 *>(ctx->d_imisc.p);
  int *d_bm = ctx->d_imisc.p + 2 * np;
  CU(cudaMemcpyAsync(ctx->d_misc.p, sx.data(), bytes_x, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_pairs, epairs.data(), bytes_p, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_bm, bm.data(), bytes_m, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_mask.p, 0, (mw + 1) * 8, ctx->stream));
  EvalParams ep;
  memset(&ep, 0, sizeof ep);
  ep.npose = 1; ep.nrb = 0; ep.nb = nb; ep.npairs = np; ep.nvc = 0; ep.object = -1;
  ep.mode_all = all_mode ? 1 : 0; ep.mask_words = mw;
  ep.fk = nullptr; ep.static_xf = reinterpret_cast<const Xf *>(ctx->d_misc.p);
  ep.body_mesh = d_bm; ep.meshes = ctx->d_meshes.p; ep.pairs = d_pairs; ep.vcs = nullptr;
  ep.legal = ctx->d_legal.p; ep.energy = nullptr; ep.pair_mask = ctx->d_mask.p;
  ep.pose_counter = ctx->d_counters.p; ep.ambig = ctx->d_ambig.p; ep.ambig_count = ctx->d_counters.p + 1;
  ep.ambig_cap = (int)ctx->d_ambig.cap; ep.stats = nullptr; ep.overflow = ctx->d_sticky.p;
  CU(ctx->d_items.reserve((size_t)np + 32));
  ep.items = ctx->d_items.p; ep.item_count = ctx->d_counters.p + 8; ep.item_cap = (int)ctx->d_items.cap;
  CU(launch_collide(ep, 1, ctx->stream)); ctx->launches += 2;
  RecheckParams rp;
  memset(&rp, 0, sizeof rp);
  rp.ambig = ctx->d_ambig.p; rp.ambig_count = ctx->d_counters.p + 1; rp.ambig_cap = (int)ctx->d_ambig.cap;
  rp.nrb = 0; rp.nb = nb; rp.fk = nullptr; rp.static_xf = ep.static_xf; rp.body_mesh = d_bm; rp.meshes = ctx->d_meshes.p;
  rp.pairs = d_pairs; rp.legal = ctx->d_legal.p; rp.energy = nullptr; rp.pair_mask = ctx->d_mask.p; rp.mask_words = mw;
  CU(launch_recheck(rp, ctx->stream)); ctx->launches++;
  uint8_t lg = 1;
  CU(cudaMemcpyAsync(&lg, ctx->d_legal.p, 1, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(mask.data(), ctx->d_mask.p, mw * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  legal = lg;
  return B2C_OK;
}

// body-body distance over explicit static bodies; one warp per query
int run_static_distance(b2c_ctx *ctx, int a, int b, double *dist, double pa[3], double pb[3]) {
  int rc = sync_meshes(ctx); if (rc) return rc;
  rc = ensure_scratch(ctx); if (rc) return rc;
  Xf sx[2] = {qt_to_xf(ctx->bodies[a].tf), qt_to_xf(ctx->bodies[b].tf)};
  int bm[2] = {ctx->bodies[a].mesh, ctx->bodies[b].mesh};
  int2 q = make_int2(0, 1);
  CU(ctx->d_misc.reserve(2 * 12 + 8));
  CU(ctx->d_imisc.reserve(16));
  CU(ctx->d_dist.reserve(4));
  CU(ctx->d_pts.reserve(8));
  int2 *d_q = reinterpret_cast<int2 *>(ctx->d_imisc.p);
  int *d_bm = ctx->d_imisc.p + 2;
  CU(cudaMemcpyAsync(ctx->d_misc.p, sx, sizeof sx, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_q, &q, sizeof q, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_bm, bm, sizeof bm, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(int), ctx->stream));
  DistParams dp;
  memset(&dp, 0, sizeof dp);
  dp.npose = 1; dp.nrb = 0; dp.nb = 2; dp.nq = 1; dp.queries = d_q; dp.fk = nullptr;
  dp.static_xf = reinterpret_cast<const Xf *>(ctx->d_misc.p); dp.body_mesh = d_bm; dp.meshes = ctx->d_meshes.p;
  dp.legal = nullptr; dp.dist = ctx->d_dist.p; dp.pts = ctx->d_pts.p; dp.work_counter = ctx->d_counters.p;
  dp.ambig = ctx->d_ambig.p; dp.ambig_count = ctx->d_counters.p + 1; dp.ambig_cap = (int)ctx->d_ambig.cap; dp.stats = nullptr; dp.overflow = ctx->d_sticky.p;
  dp.near_dir = ctx->d_near_dir.p; dp.near_ent = ctx->d_near_ent.p; dp.near_count = ctx->d_counters.p + 5;
  dp.near_cap_dir = (int)ctx->d_near_dir.cap; dp.near_cap_ent = (int)ctx->d_near_ent.cap;
  CU(launch_dist(dp, 1, ctx->stream)); ctx->launches += 2;
  RecheckParams rp;
  memset(&rp, 0, sizeof rp);
  rp.ambig = ctx->d_ambig.p; rp.ambig_count = ctx->d_counters.p + 1; rp.ambig_cap = (int)ctx->d_ambig.cap;
  rp.nrb = 0; rp.nb = 2; rp.static_xf = dp.static_xf; rp.body_mesh = d_bm; rp.meshes = ctx->d_meshes.p; rp.pairs = d_q;
  rp.dist = ctx->d_dist.p; rp.pts = ctx->d_pts.p; rp.nq = 1;
  CU(launch_recheck(rp, ctx->stream)); ctx->launches++;
  float d = 0.f;
  double pts[6];
  CU(cudaMemcpyAsync(&d, ctx->d_dist.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(pts, ctx->d_pts.p, sizeof pts, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *dist = d;
  for (int k = 0; k < 3; k++) { pa[k] = pts[k]; pb[k] = pts[3 + k]; }
  if (d >= 0.f) {
    // report the distance from the fp64 closest points (the fp32 value only steered the search)
    d3 wa = q_rot(ctx->bodies[a].tf, mk<double>(pa[0], pa[1], pa[2]));
    d3 wb = q_rot(ctx->bodies[b].tf, mk<double>(pb[0], pb[1], pb[2]));
    d3 dv = mk<double>(wa.x + ctx->bodies[a].tf.tx - wb.x - ctx->bodies[b].tf.tx,
                       wa.y + ctx->bodies[a].tf.ty - wb.y - ctx->bodies[b].tf.ty,
                       wa.z + ctx->bodies[a].tf.tz - wb.z - ctx->bodies[b].tf.tz);
    *dist = std::sqrt(dot(dv, dv));
  }
  return B2C_OK;
}

} // namespace

static void anneal_release_all(b2c_ctx *ctx);      // anneal_api.inc
static void autograsp_release_all(b2c_ctx *ctx);   // autograsp_api.inc

// =================================================================================================
extern "C" {

int b2c_create(int device, b2c_ctx **out) {
  if (!out) return B2C_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return B2C_ENODEVICE;
  b2c_ctx *ctx = new b2c_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return B2C_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (const char *e = getenv("B2C_STAGE_NODES")) ctx->stage_nodes = std::max(0, atoi(e));
  if (const char *e = getenv("B2C_SEED_RES")) ctx->seed_res = std::max(0, std::min(256, atoi(e)));
  if (const char *e = getenv("B2C_EVAL_WARPS")) ctx->eval_warps_per_block = std::max(1, std::min(32, atoi(e)));
  ctx->own_stream = ctx->stream;
  if (cudaMallocHost(&ctx->last_stats, 8 * sizeof(unsigned long long)) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return B2C_ENODEVICE; }
  memset(ctx->last_stats, 0, 8 * sizeof(unsigned long long));
  *out = ctx;
  return B2C_OK;
}

int b2c_set_stream(b2c_ctx *ctx, void *stream) {
  if (!ctx) return B2C_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
  return B2C_OK;
}

int b2c_set_profiling(b2c_ctx *ctx, int enabled) {
  if (!ctx) return B2C_EINVAL;
  cudaSetDevice(ctx->device);
  if (enabled && !ctx->ev[0])
    for (int i = 0; i < 16; i++)
      if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) return fail(ctx, B2C_ECUDA, "cudaEventCreate failed");
  ctx->profiling = enabled != 0;
  return B2C_OK;
}

int b2c_last_kernel_ms(b2c_ctx *ctx, float ms[8]) {
  if (!ctx || !ms) return B2C_EINVAL;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < 8; i++) ms[i] = 0.f;
  if (!ctx->ev[0]) return B2C_OK;
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++)
    if (ctx->ev_used[i]) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ctx->ev[2 * i], ctx->ev[2 * i + 1]) == cudaSuccess) ms[i] = t;
    }
  return B2C_OK;
}

void b2c_destroy(b2c_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();         // ctx->stream may be a caller-owned stream that no longer exists (b2c_set_stream)
  anneal_release_all(ctx);
  autograsp_release_all(ctx);
  for (int i = 0; i < 2; i++) if (ctx->h_contacts[i]) cudaFreeHost(ctx->h_contacts[i]);
  ctx->d_okeys.release(); ctx->d_odepth.release(); ctx->d_operm.release(); ctx->d_order_tab.release();
  if (ctx->h_operm) cudaFreeHost(ctx->h_operm);
  if (ctx->h_hits) cudaFreeHost(ctx->h_hits);
  if (ctx->h_nsurv) cudaFreeHost(ctx->h_nsurv);
  if (ctx->h_dtotals) cudaFreeHost(ctx->h_dtotals);
  if (ctx->h_qoff) cudaFreeHost(ctx->h_qoff);
  if (ctx->h_rout) cudaFreeHost(ctx->h_rout);
  ctx->d_rfirst.release(); ctx->d_rkeep.release(); ctx->d_rpos.release(); ctx->d_roff.release(); ctx->d_rout.release();
  if (ctx->h_qsmall) cudaFreeHost(ctx->h_qsmall);
  ctx->d_qsmall.release();
  if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
  if (ctx->ev_surv) cudaEventDestroy(ctx->ev_surv);
  ctx->d_nsurv.release(); ctx->d_soff.release(); ctx->d_dtotals.release(); ctx->d_leader.release(); ctx->d_leaders.release();
  ctx->d_items2.release();
  for (int i = 0; i < B2C_MAX_CHUNKS; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
  if (ctx->ev_keys) cudaEventDestroy(ctx->ev_keys);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->h_fk) cudaFreeHost(ctx->h_fk);
  if (ctx->last_stats) cudaFreeHost(ctx->last_stats);
  invalidate_plans(ctx);
  for (auto &m : ctx->meshes) {
    if (m.d_nodes) cudaFree(m.d_nodes);
    if (m.d_tris) cudaFree(m.d_tris);
    if (m.d_seed) cudaFree(m.d_seed);
    if (m.d_order) cudaFree(m.d_order);
    if (m.d_tri_pos) cudaFree(m.d_tri_pos);
    if (m.obb) obb_tree_free(m.obb);
  }
  ctx->d_meshes.release(); ctx->d_states.release(); ctx->d_fk.release(); ctx->d_fkqt.release(); ctx->d_legal.release();
  ctx->d_energy.release(); ctx->d_mask.release(); ctx->d_ambig.release(); ctx->d_counters.release(); ctx->d_stats.release();
  ctx->d_dist.release(); ctx->d_pts.release(); ctx->d_misc.release(); ctx->d_imisc.release(); ctx->d_legal_list.release(); ctx->d_items.release(); ctx->d_contacts.release(); ctx->d_flags.release(); ctx->d_terms.release(); ctx->d_near_dir.release(); ctx->d_near_ent.release();
  ctx->d_contacts2.release(); ctx->d_coffsets.release(); ctx->d_cubtemp.release(); ctx->d_sticky.release();
  for (int i = 0; i < 16; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char *b2c_last_error(const b2c_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *b2c_stream(const b2c_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int b2c_device(const b2c_ctx *ctx) { return ctx ? ctx->device : -1; }
long long b2c_launch_count(const b2c_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- geometry ------------------------------------------------------------------------------------
int b2c_mesh_create(b2c_ctx *ctx, const float *tri9, int ntri, int *mesh_id) {
  if (!ctx || !mesh_id || ntri < 0 || (ntri > 0 && !tri9)) return fail(ctx, B2C_EINVAL, "mesh_create: bad arguments");
  if (ntri >= (1 << 23)) return fail(ctx, B2C_ECAPACITY, "mesh_create: %d triangles exceed the 2^23 per-mesh limit", ntri);
  cudaSetDevice(ctx->device);
  MeshHost m;
  m.alive = true;
  m.ntri = ntri;
  m.tri9.assign(tri9, tri9 + 9 * (size_t)ntri);
  Node r;
  build_bvh(m.tri9.data(), ntri, m.nodes, &m.quant, &m.depth, &r);
  m.extent = ntri > 0 ? std::fabs(r.cx) + std::fabs(r.cy) + std::fabs(r.cz) + r.hx + r.hy + r.hz : 0.f;
  m.root_c[0] = r.cx; m.root_c[1] = r.cy; m.root_c[2] = r.cz; m.root_h[0] = r.hx; m.root_h[1] = r.hy; m.root_h[2] = r.hz;
  std::vector<Tri> tris(std::max(ntri, 1));
  // an empty mesh is one far-away degenerate triangle (its leaf box sits at the same place): every query answers
  // "no collision, nothing near" without special cases in the kernels
  if (ntri == 0) for (int v = 0; v < 3; v++) tris[0].v[v] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);
  for (int i = 0; i < ntri; i++)
    for (int v = 0; v < 3; v++)
      tris[i].v[v] = make_float4(tri9[9 * (size_t)i + 3 * v], tri9[9 * (size_t)i + 3 * v + 1], tri9[9 * (size_t)i + 3 * v + 2], 0.f);
  // v[0].w = radius of the triangle about its first vertex, rounded up: region_batch_kernel's gate bounds a triangle's
  // distance to a point from that one 16-byte load
  for (int i = 0; i < ntri; i++) {
    const float *t = tri9 + 9 * (size_t)i;
    double e1 = 0.0, e2 = 0.0;
    for (int k = 0; k < 3; k++) {
      e1 += ((double)t[3 + k] - (double)t[k]) * ((double)t[3 + k] - (double)t[k]);
      e2 += ((double)t[6 + k] - (double)t[k]) * ((double)t[6 + k] - (double)t[k]);
    }
    tris[i].v[0].w = (float)(std::sqrt(std::max(e1, e2)) * 1.00001 + 1.0e-30);
  }
  CU(cudaMalloc(&m.d_nodes, m.nodes.size() * sizeof(NodeQ)));
  CU(cudaMalloc(&m.d_tris, tris.size() * sizeof(Tri)));
  CU(cudaMemcpy(m.d_nodes, m.nodes.data(), m.nodes.size() * sizeof(NodeQ), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(m.d_tris, tris.data(), tris.size() * sizeof(Tri), cudaMemcpyHostToDevice));
  ctx->meshes.push_back(std::move(m));
  ctx->meshes_dirty = true;
  *mesh_id = (int)ctx->meshes.size() - 1;
  return B2C_OK;
}

int b2c_mesh_destroy(b2c_ctx *ctx, int mesh_id) {
  if (!ctx || !mesh_ok(ctx, mesh_id)) return fail(ctx, B2C_EINVAL, "mesh_destroy: unknown mesh %d", mesh_id);
  for (auto &b : ctx->bodies)
    if (b.alive && b.mesh == mesh_id) return fail(ctx, B2C_EINVAL, "mesh_destroy: mesh %d still used by a body", mesh_id);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  MeshHost &m = ctx->meshes[mesh_id];
  cudaFree(m.d_nodes); cudaFree(m.d_tris);
  if (m.d_seed) cudaFree(m.d_seed);
  if (m.d_order) cudaFree(m.d_order);
  if (m.d_tri_pos) cudaFree(m.d_tri_pos);
  if (m.obb) obb_tree_free(m.obb);
  m = MeshHost();
  ctx->meshes_dirty = true;
  return B2C_OK;
}

int b2c_body_create(b2c_ctx *ctx, int mesh_id, int *body_id) {
  if (!ctx || !body_id || !mesh_ok(ctx, mesh_id)) return fail(ctx, B2C_EINVAL, "body_create: unknown mesh %d", mesh_id);
  BodyHost b;
  b.alive = true; b.mesh = mesh_id; b.active = true; b.tf = q_identity();
  ctx->bodies.push_back(b);
  *body_id = (int)ctx->bodies.size() - 1;
  invalidate_plans(ctx);
  return B2C_OK;
}

int b2c_body_set_mesh(b2c_ctx *ctx, int body_id, int mesh_id) {
  if (!ctx || !body_ok(ctx, body_id) || !mesh_ok(ctx, mesh_id)) return fail(ctx, B2C_EINVAL, "body_set_mesh: unknown body/mesh");
  ctx->bodies[body_id].mesh = mesh_id;
  invalidate_plans(ctx);
  return B2C_OK;
}

int b2c_body_remove(b2c_ctx *ctx, int body_id) {
  if (!ctx || !body_ok(ctx, body_id)) return fail(ctx, B2C_EINVAL, "body_remove: unknown body %d", body_id);
  ctx->bodies[body_id].alive = false;
  for (auto it = ctx->disabled.begin(); it != ctx->disabled.end();)
    if (it->first == body_id || it->second == body_id) it = ctx->disabled.erase(it); else ++it;
  invalidate_plans(ctx);
  return B2C_OK;
}

int b2c_body_set_transform(b2c_ctx *ctx, int body_id, const double q[4], const double t[3]) {
  if (!ctx || !body_ok(ctx, body_id) || !q || !t) return fail(ctx, B2C_EINVAL, "set_transform: unknown body %d", body_id);
  ctx->bodies[body_id].tf = make_qt(q, t);
  return B2C_OK;
}

int b2c_body_get_transform(b2c_ctx *ctx, int body_id, double q[4], double t[3]) {
  if (!ctx || !body_ok(ctx, body_id) || !q || !t) return fail(ctx, B2C_EINVAL, "get_transform: unknown body %d", body_id);
  const Qt &f = ctx->bodies[body_id].tf;
  q[0] = f.w; q[1] = f.x; q[2] = f.y; q[3] = f.z; t[0] = f.tx; t[1] = f.ty; t[2] = f.tz;
  return B2C_OK;
}

int b2c_body_set_active(b2c_ctx *ctx, int body_id, int active) {
  if (!ctx || !body_ok(ctx, body_id)) return fail(ctx, B2C_EINVAL, "set_active: unknown body %d", body_id);
  ctx->bodies[body_id].active = active != 0;
  invalidate_plans(ctx);
  return B2C_OK;
}

/* CollisionInterface::newThread / getThreadId (collisionInterface.h:181-184) and CollisionModel's thread id */
int b2c_body_set_thread(b2c_ctx *ctx, int body_id, int thread_id) {
  if (!ctx || !body_ok(ctx, body_id) || thread_id < 0) return fail(ctx, B2C_EINVAL, "body_set_thread: unknown body %d", body_id);
  ctx->bodies[body_id].thread = thread_id;
  invalidate_plans(ctx);
  return B2C_OK;
}
int b2c_set_thread(b2c_ctx *ctx, int thread_id) {
  if (!ctx || thread_id < 0) return fail(ctx, B2C_EINVAL, "set_thread: bad thread id");
  if (ctx->query_thread != thread_id) { ctx->query_thread = thread_id; invalidate_plans(ctx); }
  return B2C_OK;
}

int b2c_pair_set_active(b2c_ctx *ctx, int a, int b, int active) {
  if (!ctx || !body_ok(ctx, a) || !body_ok(ctx, b)) return fail(ctx, B2C_EINVAL, "pair_set_active: unknown body");
  if (a == b) {   // reference: "insertion collision pair is actually one body" -> toggles the body
    ctx->bodies[a].active = active != 0;
  } else {
    std::pair<int, int> k(std::min(a, b), std::max(a, b));
    if (active) ctx->disabled.erase(k); else ctx->disabled.insert(k);
  }
  invalidate_plans(ctx);
  return B2C_OK;
}

int b2c_pair_is_active(b2c_ctx *ctx, int a, int b, int *active) {
  if (!ctx || !active || !body_ok(ctx, a)) return fail(ctx, B2C_EINVAL, "pair_is_active: unknown body");
  if (b < 0 || a == b) { *active = ctx->bodies[a].active ? 1 : 0; return B2C_OK; }
  if (!body_ok(ctx, b)) return fail(ctx, B2C_EINVAL, "pair_is_active: unknown body");
  *active = pair_active(ctx, a, b) ? 1 : 0;
  return B2C_OK;
}

// ---- single-pose queries ---------------------------------------------------------------------------
int b2c_all_collisions(b2c_ctx *ctx, int fast, const int *interest, int ninterest, int *pairs_out, int cap, int *ncolliding) {
  if (!ctx || !ncolliding) return fail(ctx, B2C_EINVAL, "all_collisions: bad arguments");
  cudaSetDevice(ctx->device);
  std::set<int> il;
  if (interest) for (int i = 0; i < ninterest; i++) if (body_ok(ctx, interest[i])) il.insert(interest[i]);
  std::vector<std::pair<int, int>> pairs;
  active_pairs(ctx, interest ? &il : nullptr, pairs);
  std::vector<unsigned long long> mask;
  int legal = 1;
  int rc = run_static_collisions(ctx, pairs, !fast, mask, legal);
  if (rc) return rc;
  if (fast) { *ncolliding = legal ? 0 : 1; return B2C_OK; }
  int n = 0;
  for (size_t i = 0; i < pairs.size(); i++)
    if ((mask[i >> 6] >> (i & 63)) & 1ull) {
      if (pairs_out && n < cap) { pairs_out[2 * n] = pairs[i].first; pairs_out[2 * n + 1] = pairs[i].second; }
      n++;
    }
  *ncolliding = n;
  return B2C_OK;
}

int b2c_body_distance(b2c_ctx *ctx, int a, int b, double *dist, double p1[3], double p2[3], double p1l[3], double p2l[3]) {
  if (!ctx || !dist) return fail(ctx, B2C_EINVAL, "body_distance: bad arguments");
  if (!body_ok(ctx, a) || !body_ok(ctx, b)) { *dist = 0.0; return fail(ctx, B2C_EINVAL, "body_distance: model not found"); }
  cudaSetDevice(ctx->device);
  double pa[3], pb[3];
  int rc = run_static_distance(ctx, a, b, dist, pa, pb);
  if (rc) return rc;
  if (p1l) for (int k = 0; k < 3; k++) p1l[k] = pa[k];
  if (p2l) for (int k = 0; k < 3; k++) p2l[k] = pb[k];
  // reference quirk: the already-local points are pushed through the inverse body transforms once more
  // (graspitCollision.cpp:473-475)
  if (p1) { Qt inv = q_inverse(ctx->bodies[a].tf); d3 r = q_rot(inv, mk<double>(pa[0], pa[1], pa[2])); p1[0] = r.x + inv.tx; p1[1] = r.y + inv.ty; p1[2] = r.z + inv.tz; }
  if (p2) { Qt inv = q_inverse(ctx->bodies[b].tf); d3 r = q_rot(inv, mk<double>(pb[0], pb[1], pb[2])); p2[0] = r.x + inv.tx; p2[1] = r.y + inv.ty; p2[2] = r.z + inv.tz; }
  return B2C_OK;
}

int b2c_point_distance(b2c_ctx *ctx, int body, const double point[3], double *dist, double closest[3], double normal[3]) {
  if (!ctx || !dist || !point) return fail(ctx, B2C_EINVAL, "point_distance: bad arguments");
  if (!body_ok(ctx, body)) { *dist = 0.0; return fail(ctx, B2C_EINVAL, "point_distance: model not found"); }
  cudaSetDevice(ctx->device);
  int rc = sync_meshes(ctx); if (rc) return rc;
  CU(ctx->d_misc.reserve(12 + 3 + 7 + 8));
  Xf x = qt_to_xf(ctx->bodies[body].tf);
  double *d_x = ctx->d_misc.p, *d_p = d_x + 12, *d_o = d_p + 4;
  CU(cudaMemcpyAsync(d_x, &x, sizeof x, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_p, point, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  PointParams pp;
  pp.n = 1; pp.points = d_p; pp.body_xf = reinterpret_cast<const Xf *>(d_x);
  const MeshHost &m = ctx->meshes[ctx->bodies[body].mesh];
  pp.mesh.nodes = m.d_nodes; pp.mesh.tris = m.d_tris; pp.mesh.nnodes = (int)m.nodes.size(); pp.mesh.ntri = m.ntri;
  for (int k = 0; k < 3; k++) { pp.mesh.q0[k] = m.quant.q0[k]; pp.mesh.qs[k] = m.quant.qs[k]; }
  pp.mesh.extent = m.extent; pp.mesh.depth = m.depth; pp.mesh.seed = nullptr;
  pp.out = d_o;
  CU(launch_point(pp, ctx->stream)); ctx->launches++;
  double o[7];
  CU(cudaMemcpyAsync(o, d_o, sizeof o, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *dist = o[0];
  if (closest) for (int k = 0; k < 3; k++) closest[k] = o[1 + k];
  if (normal) for (int k = 0; k < 3; k++) normal[k] = o[4 + k];
  return B2C_OK;
}

int b2c_bounding_volumes(b2c_ctx *ctx, int body, int depth, b2c_obb *out, int cap, int *nboxes) {
  if (!ctx || !nboxes) return fail(ctx, B2C_EINVAL, "bounding_volumes: bad arguments");
  if (!body_ok(ctx, body)) { *nboxes = 0; return fail(ctx, B2C_EINVAL, "bounding_volumes: model not found"); }
  MeshHost &m = ctx->meshes[ctx->bodies[body].mesh];
  if (!m.obb) m.obb = obb_tree_build(m.tri9.data(), m.ntri);
  std::vector<ObbBox> boxes;
  obb_tree_level(m.obb, depth, boxes);
  *nboxes = (int)boxes.size();
  for (int i = 0; i < (int)boxes.size() && i < cap && out; i++) {
    for (int k = 0; k < 3; k++) { out[i].t[k] = boxes[i].t[k]; out[i].half[k] = boxes[i].half[k]; }
    for (int k = 0; k < 4; k++) out[i].q_wxyz[k] = boxes[i].q[k];
  }
  return B2C_OK;
}

// ---- robots ----------------------------------------------------------------------------------------
int b2c_robot_define(b2c_ctx *ctx, const b2c_robot_desc *d, int *robot_id) {
  if (!ctx || !d || !robot_id) return fail(ctx, B2C_EINVAL, "robot_define: bad arguments");
  if (!body_ok(ctx, d->base_body)) return fail(ctx, B2C_EINVAL, "robot_define: unknown base body %d", d->base_body);
  if (d->ndof > FK_MAX_DOF_HOST) return fail(ctx, B2C_ECAPACITY, "robot_define: more than %d DOFs", FK_MAX_DOF_HOST);
  if (d->ndof < 0 || d->nchains < 0 || d->njoints < 0 || d->nlinks < 0 || (d->ndof > 0 && (!d->dof_min || !d->dof_max)) ||
      (d->nchains > 0 && !d->chains) || (d->njoints > 0 && !d->joints) || (d->nlinks > 0 && (!d->link_body || !d->link_last_joint)))
    return fail(ctx, B2C_EINVAL, "robot_define: inconsistent counts / null tables");
  if (d->mount_body >= 0 && !body_ok(ctx, d->mount_body)) return fail(ctx, B2C_EINVAL, "robot_define: unknown mount body %d", d->mount_body);
  for (int c = 0; c < d->nchains; c++) {
    const b2c_chain_desc &cd = d->chains[c];
    if (cd.first_joint < 0 || cd.njoints < 0 || cd.first_joint + cd.njoints > d->njoints || cd.first_link < 0 || cd.nlinks < 1 ||
        cd.first_link + cd.nlinks > d->nlinks)
      return fail(ctx, B2C_EINVAL, "robot_define: chain %d: joint / link range outside the tables", c);
    for (int l = 0; l < cd.nlinks; l++) {
      const int lj = d->link_last_joint[cd.first_link + l];
      if (lj < 0 || lj >= cd.njoints) return fail(ctx, B2C_EINVAL, "robot_define: chain %d link %d: last joint %d outside the chain", c, l, lj);
    }
  }
  for (int j = 0; j < d->njoints; j++)
    if (d->joints[j].dof < 0 || d->joints[j].dof >= d->ndof) return fail(ctx, B2C_EINVAL, "robot_define: joint %d: DOF %d outside [0, %d)", j, d->joints[j].dof, d->ndof);
  if (d->parent_robot >= 0) {
    if (d->parent_robot >= (int)ctx->robots.size()) return fail(ctx, B2C_EINVAL, "robot_define: parent robot %d not defined yet", d->parent_robot);
    if (d->parent_chain < 0 || d->parent_chain >= ctx->robots[d->parent_robot].d.nchains)
      return fail(ctx, B2C_EINVAL, "robot_define: parent chain %d outside the parent robot's %d chains", d->parent_chain, ctx->robots[d->parent_robot].d.nchains);
  }
  RobotHost r;
  r.d = *d;
  r.dof_min.assign(d->dof_min, d->dof_min + d->ndof);
  r.dof_max.assign(d->dof_max, d->dof_max + d->ndof);
  r.chains.assign(d->chains, d->chains + d->nchains);
  r.joints.assign(d->joints, d->joints + d->njoints);
  r.link_body.assign(d->link_body, d->link_body + d->nlinks);
  r.link_last_joint.assign(d->link_last_joint, d->link_last_joint + d->nlinks);
  for (int b : r.link_body) if (!body_ok(ctx, b)) return fail(ctx, B2C_EINVAL, "robot_define: unknown link body %d", b);
  if (d->neg > 0 && d->eg) {
    r.eg.assign(d->eg, d->eg + (size_t)d->neg * d->ndof);
    r.eg_origin.assign(d->eg_origin, d->eg_origin + d->ndof);
    r.eg_norm.assign(d->eg_norm, d->eg_norm + d->ndof);
  } else {
    r.d.neg = 0;
  }
  if (d->nvcontacts > 0) r.vcs.assign(d->vcontacts, d->vcontacts + d->nvcontacts);
  if (d->parent_robot >= (int)ctx->robots.size()) return fail(ctx, B2C_EINVAL, "robot_define: parent robot %d not defined yet", d->parent_robot);
  ctx->robots.push_back(std::move(r));
  *robot_id = (int)ctx->robots.size() - 1;
  invalidate_plans(ctx);
  return B2C_OK;
}

} // extern "C"

namespace {

void collect_robots(const b2c_ctx *c, int root, std::vector<int> &order) {
  order.push_back(root);
  for (int r = 0; r < (int)c->robots.size(); r++)
    if (c->robots[r].d.parent_robot == root) collect_robots(c, r, order);
}

// bodies of one robot in World::findVirtualContacts order: links chain-major, base, mount
void robot_own_bodies(const RobotHost &r, std::vector<int> &out) {
  for (int b : r.link_body) out.push_back(b);
  out.push_back(r.d.base_body);
  if (r.d.mount_body >= 0) out.push_back(r.d.mount_body);
}

// floats (doubles for JOINTS) one pose occupies at least, per input mode
int input_min_stride(int mode, int ndof_total, int njoints_total, int ndof_root, int neg) {
  switch (mode) {
    case B2C_INPUT_RAW: return 7 + ndof_total;
    case B2C_INPUT_AA_EIGEN: return 6 + neg;
    case B2C_INPUT_JOINTS: return 7 + njoints_total;
    case B2C_INPUT_ELLIPSOID_EIGEN: return 4 + neg;
    case B2C_INPUT_APPROACH_EIGEN: return 3 + neg;
    default: return 7 + ndof_root;      // B2C_INPUT_COMPLETE_DOF
  }
}
bool input_needs_eigen(int mode) { return mode == B2C_INPUT_AA_EIGEN || mode == B2C_INPUT_ELLIPSOID_EIGEN || mode == B2C_INPUT_APPROACH_EIGEN; }

int build_plan(b2c_ctx *ctx, int robot, int object, Plan **out) {
  auto key = std::make_pair(robot, object);
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end() && it->second->valid) { *out = it->second; return B2C_OK; }
  Plan *P = new Plan();
  P->robot = robot; P->object = object;
  collect_robots(ctx, robot, P->robots);
  if ((int)P->robots.size() > FK_MAX_ROBOTS) { delete P; return fail(ctx, B2C_ECAPACITY, "more than %d robots in one kinematic tree", FK_MAX_ROBOTS); }
  for (int rr : P->robots) {
    std::vector<int> own;
    robot_own_bodies(ctx->robots[rr], own);
    for (int b : own)
      if (!body_ok(ctx, b)) { delete P; return fail(ctx, B2C_EINVAL, "robot %d refers to body %d, which was removed", rr, b); }
  }
  std::map<int, int> slot_of;
  for (size_t i = 0; i < P->robots.size(); i++) {
    std::vector<int> own;
    robot_own_bodies(ctx->robots[P->robots[i]], own);
    // LIVE contacts and link distances cover the finger links and the palm of the hand only -- not the mount piece
    // (World::findVirtualContacts, reference world.cpp:1619-1639); they are the first nlinks + 1 slots
    if (i == 0) P->nhand = (int)ctx->robots[P->robots[i]].link_body.size() + 1;
    for (int b : own) { slot_of[b] = (int)P->slots.size(); P->slots.push_back(b); }
  }
  // FK program
  std::vector<FkRobot> fr; std::vector<FkChain> fc; std::vector<FkJoint> fj; std::vector<FkLink> fl; std::vector<double> eg;
  int dof_ofs = 0;
  std::map<int, int> prog_index;
  for (size_t i = 0; i < P->robots.size(); i++) {
    const RobotHost &r = ctx->robots[P->robots[i]];
    prog_index[P->robots[i]] = (int)i;
    FkRobot R;
    memset(&R, 0, sizeof R);
    R.base_slot = slot_of[r.d.base_body];
    R.mount_slot = r.d.mount_body >= 0 ? slot_of[r.d.mount_body] : -1;
    R.parent = -1; R.parent_last_slot = -1; R.parent_offset = q_identity();
    if (i > 0) {
      const RobotHost &pr = ctx->robots[r.d.parent_robot];
      R.parent = prog_index[r.d.parent_robot];
      const b2c_chain_desc &pc = pr.chains[r.d.parent_chain];
      R.parent_last_slot = slot_of[pr.link_body[pc.first_link + pc.nlinks - 1]];
      R.parent_offset = make_qt(r.d.parent_offset_q, r.d.parent_offset_t);
    }
    R.dof_ofs = dof_ofs; R.ndof = r.d.ndof;
    R.first_chain = (int)fc.size(); R.nchains = r.d.nchains;
    R.approach = make_qt(r.d.approach_q, r.d.approach_t);
    R.approach_inv = q_inverse(R.approach);
    R.neg = r.d.neg;
    R.eg_ofs = (int)eg.size();
    if (r.d.neg > 0) {
      eg.insert(eg.end(), r.eg.begin(), r.eg.end());
      eg.insert(eg.end(), r.eg_origin.begin(), r.eg_origin.end());
      eg.insert(eg.end(), r.eg_norm.begin(), r.eg_norm.end());
      eg.insert(eg.end(), r.dof_min.begin(), r.dof_min.end());
      eg.insert(eg.end(), r.dof_max.begin(), r.dof_max.end());
    }
    for (int ci = 0; ci < r.d.nchains; ci++) {
      const b2c_chain_desc &c = r.chains[ci];
      FkChain C;
      C.tran = make_qt(c.tran_q, c.tran_t);
      C.first_joint = (int)fj.size(); C.njoints = c.njoints;
      C.first_link = (int)fl.size(); C.nlinks = c.nlinks;
      for (int j = 0; j < c.njoints; j++) {
        const b2c_joint_desc &jd = r.joints[c.first_joint + j];
        FkJoint J;
        J.dof = dof_ofs + jd.dof; J.revolute = jd.revolute; J.ratio = jd.ratio; J.theta = jd.theta; J.d = jd.d;
        Qt tr3 = q_identity(); tr3.tx = jd.a;
        J.tr4tr3 = q_compose(tr3, q_axis_angle(jd.alpha, 1.0, 0.0, 0.0));
        fj.push_back(J);
      }
      for (int l = 0; l < c.nlinks; l++) {
        FkLink L;
        L.out_slot = slot_of[r.link_body[c.first_link + l]];
        L.last_joint = r.link_last_joint[c.first_link + l];
        fl.push_back(L);
      }
      fc.push_back(C);
    }
    dof_ofs += r.d.ndof;
    fr.push_back(R);
  }
  P->ndof_total = dof_ofs;
  if (dof_ofs > FK_MAX_DOF_HOST) { delete P; return fail(ctx, B2C_ECAPACITY, "more than %d DOFs in the robot tree", FK_MAX_DOF_HOST); }

  // eval bodies and pairs: every active pair touching a robot body (World::noCollision interest list,
  // src/world.cpp:1491-1503)
  P->ebodies = P->slots;
  std::map<int, int> eidx;
  for (size_t i = 0; i < P->slots.size(); i++) eidx[P->slots[i]] = (int)i;
  std::set<int> interest(P->slots.begin(), P->slots.end());
  std::vector<std::pair<int, int>> ap;
  active_pairs(ctx, &interest, ap);
  auto eb = [&](int b) {
    auto f = eidx.find(b);
    if (f != eidx.end()) return f->second;
    eidx[b] = (int)P->ebodies.size();
    P->ebodies.push_back(b);
    return eidx[b];
  };
  for (auto &pr : ap) {
    P->pairs.push_back(make_int2(eb(pr.first), eb(pr.second)));
    P->pair_ids.push_back(make_int2(pr.first, pr.second));
  }
  P->object_e = object >= 0 ? eb(object) : -1;
  // virtual contacts of the root robot (Grasp::collectVirtualContacts order is irrelevant to the mean)
  // (for an arm with a hand attached the contacts are the hand's: every robot of the tree contributes its own)
  std::vector<VcDev> vcs;
  for (int rr : P->robots)
    for (auto &v : ctx->robots[rr].vcs) {
      VcDev d;
      memset(&d, 0, sizeof d);
      auto f = eidx.find(v.body);
      if (f == eidx.end()) continue;
      d.body = f->second;
      for (int k = 0; k < 3; k++) { d.loc[k] = v.loc[k]; d.normal[k] = v.normal[k]; }
      vcs.push_back(d);
    }
  P->nvc = (int)vcs.size();
  std::vector<int> bm(P->ebodies.size());
  for (size_t i = 0; i < bm.size(); i++) bm[i] = ctx->bodies[P->ebodies[i]].mesh;
  std::vector<int2> queries;
  if (P->object_e >= 0) for (int i = 0; i < P->nhand; i++) queries.push_back(make_int2(i, P->object_e));

#define UP(buf, vec)                                                                                        \
  do {                                                                                                      \
    if (P->buf.reserve((vec).size() + 1) != cudaSuccess) { delete P; return fail(ctx, B2C_ECUDA, "plan upload: cudaMalloc failed"); } \
    if (!(vec).empty() && cudaMemcpy(P->buf.p, (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice) != cudaSuccess) { \
      delete P; return fail(ctx, B2C_ECUDA, "plan upload: cudaMemcpy failed"); }                            \
  } while (0)
  UP(d_robots, fr); UP(d_chains, fc); UP(d_joints, fj); UP(d_links, fl); UP(d_eg, eg);
  UP(d_body_mesh, bm); UP(d_pairs, P->pairs); UP(d_vcs, vcs); UP(d_queries, queries);
#undef UP
  if (P->d_static.reserve(P->ebodies.size() - P->slots.size() + 1) != cudaSuccess) { delete P; return fail(ctx, B2C_ECUDA, "plan upload: cudaMalloc failed"); }
  P->prog.robots = P->d_robots.p; P->prog.chains = P->d_chains.p; P->prog.joints = P->d_joints.p; P->prog.links = P->d_links.p;
  P->prog.egdata = P->d_eg.p; P->prog.nrobots = (int)fr.size(); P->prog.nslots = (int)P->slots.size(); P->prog.ndof_total = dof_ofs;
  P->valid = true;
  ctx->plans[key] = P;
  *out = P;
  return B2C_OK;
}

} // namespace

extern "C" {

int b2c_robot_bodies(b2c_ctx *ctx, int robot, int *bodies, int cap, int *nbodies) {
  if (!ctx || !nbodies || robot < 0 || robot >= (int)ctx->robots.size()) return fail(ctx, B2C_EINVAL, "robot_bodies: unknown robot");
  cudaSetDevice(ctx->device);
  Plan *P;
  int rc = build_plan(ctx, robot, -1, &P);
  if (rc) return rc;
  *nbodies = (int)P->slots.size();
  for (int i = 0; i < (int)P->slots.size() && i < cap && bodies; i++) bodies[i] = P->slots[i];
  return B2C_OK;
}

int b2c_robot_pairs(b2c_ctx *ctx, int robot, int *pairs, int cap, int *npairs) {
  if (!ctx || !npairs || robot < 0 || robot >= (int)ctx->robots.size()) return fail(ctx, B2C_EINVAL, "robot_pairs: unknown robot");
  cudaSetDevice(ctx->device);
  Plan *P;
  int rc = build_plan(ctx, robot, -1, &P);
  if (rc) return rc;
  *npairs = (int)P->pair_ids.size();
  for (int i = 0; i < (int)P->pair_ids.size() && i < cap && pairs; i++) { pairs[2 * i] = P->pair_ids[i].x; pairs[2 * i + 1] = P->pair_ids[i].y; }
  return B2C_OK;
}

int b2c_last_stats(b2c_ctx *ctx, unsigned long long stats[8]) {
  if (!ctx || !stats) return B2C_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);       // the counters of the last call ride on its stream
  for (int i = 0; i < 8; i++) stats[i] = ctx->last_stats[i];
  return B2C_OK;
}

int b2c_eval_batch(b2c_ctx *ctx, const b2c_eval_args *a) {
  if (!ctx || !a) return fail(ctx, B2C_EINVAL, "eval_batch: null arguments");
  if (a->robot < 0 || a->robot >= (int)ctx->robots.size()) return fail(ctx, B2C_EINVAL, "eval_batch: unknown robot %d", a->robot);
  if (a->object >= 0 && !body_ok(ctx, a->object)) return fail(ctx, B2C_EINVAL, "eval_batch: unknown object body %d", a->object);
  if (a->npose < 0 || !a->states) return fail(ctx, B2C_EINVAL, "eval_batch: bad pose batch");
  if (a->npose == 0) return B2C_OK;
  cudaSetDevice(ctx->device);
  int rc = B2C_OK;
  if (a->object >= 0) { rc = ensure_seed_grid(ctx, ctx->bodies[a->object].mesh); if (rc) return rc; }
  rc = sync_meshes(ctx); if (rc) return rc;
  rc = ensure_scratch(ctx); if (rc) return rc;
  Plan *P;
  rc = build_plan(ctx, a->robot, a->object, &P); if (rc) return rc;
  const RobotHost &root = ctx->robots[a->robot];
  const int npose = a->npose;
#define PROF_BEGIN(i) do { if (ctx->profiling) { cudaEventRecord(ctx->ev[2 * (i)], ctx->stream); ctx->ev_used[i] = true; } } while (0)
#define PROF_END(i) do { if (ctx->profiling) cudaEventRecord(ctx->ev[2 * (i) + 1], ctx->stream); } while (0)
  for (int i = 0; i < 8; i++) ctx->ev_used[i] = false;
  const bool dev = (a->flags & B2C_EVAL_DEVICE_PTRS) != 0;
  const bool live = (a->flags & B2C_EVAL_LIVE) != 0;
  const bool want_mask = (a->flags & B2C_EVAL_PAIRMASK) != 0 && a->pair_mask;
  const bool want_dist = (a->flags & B2C_EVAL_LINKDIST) != 0 && a->link_dist;
  const bool joints_in = a->input_mode == B2C_INPUT_JOINTS;
  int njoints_total = 0;
  for (int r : P->robots) njoints_total += ctx->robots[r].d.njoints;
  if (a->input_mode < 0 || a->input_mode > B2C_INPUT_COMPLETE_DOF) return fail(ctx, B2C_EINVAL, "eval_batch: unknown input mode %d", a->input_mode);
  const int min_stride = input_min_stride(a->input_mode, P->ndof_total, njoints_total, root.d.ndof, root.d.neg);
  if (a->stride < min_stride) return fail(ctx, B2C_EINVAL, "eval_batch: stride %d < %d", a->stride, min_stride);
  if (input_needs_eigen(a->input_mode) && root.d.neg <= 0) return fail(ctx, B2C_EINVAL, "eval_batch: robot has no eigengrasps");
  if ((live || want_dist) && a->object < 0) return fail(ctx, B2C_EINVAL, "eval_batch: LIVE / LINKDIST need an object");
  if ((live || want_dist) && P->robots.size() > 1)
    return fail(ctx, B2C_EINVAL, "eval_batch: LIVE / LINKDIST are defined over the hand's own links and palm; evaluate the hand robot, not the arm it is attached to");
  const int nrb = (int)P->slots.size(), nb = (int)P->ebodies.size(), np = (int)P->pairs.size();
  const int mw = std::max(1, (np + 63) / 64);

  // inputs
  const float *d_states = a->states;
  if (!dev) {
    const size_t elem = joints_in ? sizeof(double) : sizeof(float);
    CU(ctx->d_states.reserve((size_t)npose * a->stride * (elem / sizeof(float))));
    CU(cudaMemcpyAsync(ctx->d_states.p, a->states, (size_t)npose * a->stride * elem, cudaMemcpyHostToDevice, ctx->stream));
    d_states = ctx->d_states.p;
  }
  CU(ctx->d_fk.reserve((size_t)npose * nrb));
  const bool want_tf = a->body_tf != nullptr;
  if (want_tf && !dev) CU(ctx->d_fkqt.reserve((size_t)npose * nrb * 7));
  uint8_t *d_legal = dev && a->legal ? a->legal : nullptr;
  float *d_energy = dev && a->energy ? a->energy : nullptr;
  unsigned long long *d_mask = dev && want_mask ? (unsigned long long *)a->pair_mask : nullptr;
  if (!d_legal) { CU(ctx->d_legal.reserve(npose)); d_legal = ctx->d_legal.p; }
  if (!d_energy) { CU(ctx->d_energy.reserve(npose)); d_energy = ctx->d_energy.p; }
  if (want_mask && !d_mask) { CU(ctx->d_mask.reserve((size_t)npose * mw)); d_mask = ctx->d_mask.p; }
  // static transforms
  std::vector<Xf> sx(nb - nrb);
  for (int i = nrb; i < nb; i++) sx[i - nrb] = qt_to_xf(ctx->bodies[P->ebodies[i]].tf);
  // uploaded only when a static body moved since the last call: a pageable H2D copy synchronises the stream first, which
  // would tie the host to the GPU once per annealing step
  if (!sx.empty() && (P->static_cache.size() != sx.size() || memcmp(P->static_cache.data(), sx.data(), sx.size() * sizeof(Xf)) != 0)) {
    CU(cudaMemcpyAsync(P->d_static.p, sx.data(), sx.size() * sizeof(Xf), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    P->static_cache = sx;
  }
  CU(cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_stats.p, 0, 8 * sizeof(unsigned long long), ctx->stream));

  // K1
  FkParams fp;
  memset(&fp, 0, sizeof fp);
  fp.prog = P->prog; fp.npose = npose; fp.input_mode = a->input_mode; fp.stride = a->stride; fp.states = d_states;
  fp.jstates = joints_in ? reinterpret_cast<const double *>(d_states) : nullptr;
  fp.ref = make_qt(a->ref_q, a->ref_t);
  {
    // PositionStateEllipsoid's parameters (searchStateImpl.cpp:212-215); all zero = the reference's defaults
    const bool unset = a->position_params[0] == 0.0 && a->position_params[1] == 0.0 && a->position_params[2] == 0.0;
    fp.pos_params[0] = unset ? 80.0 : a->position_params[0];
    fp.pos_params[1] = unset ? 80.0 : a->position_params[1];
    fp.pos_params[2] = unset ? 160.0 : a->position_params[2];
  }
  fp.out = ctx->d_fk.p;
  fp.out_qt = want_tf ? (dev ? a->body_tf : ctx->d_fkqt.p) : nullptr;
  PROF_BEGIN(0); CU(launch_fk(fp, ctx->stream)); ctx->launches++; PROF_END(0);

  // K3 (+K5+K8 when PRESET)
  EvalParams ep;
  memset(&ep, 0, sizeof ep);
  ep.npose = npose; ep.nrb = nrb; ep.nb = nb; ep.npairs = np; ep.nvc = P->nvc; ep.object = P->object_e;
  ep.mode_all = want_mask ? 1 : 0;
  ep.mask_words = mw;
  ep.fk = ctx->d_fk.p; ep.static_xf = P->d_static.p; ep.body_mesh = P->d_body_mesh.p; ep.meshes = ctx->d_meshes.p;
  ep.pairs = P->d_pairs.p; ep.vcs = P->d_vcs.p; ep.legal = d_legal; ep.energy = d_energy; ep.pair_mask = d_mask;
  ep.pose_counter = ctx->d_counters.p; ep.ambig = ctx->d_ambig.p; ep.ambig_count = ctx->d_counters.p + 1;
  ep.ambig_cap = (int)ctx->d_ambig.cap; ep.stats = ctx->d_stats.p; ep.overflow = ctx->d_sticky.p;
  const bool want_energy = a->object >= 0 && (live || P->nvc > 0);
  if (want_energy) {
    CU(ctx->d_legal_list.reserve(npose));
    ep.legal_list = ctx->d_legal_list.p; ep.legal_count = ctx->d_counters.p + 4;
  }
  {
    size_t want = (size_t)npose * np + 32;
    if (want > (size_t)0x7fffffff) return fail(ctx, B2C_ECAPACITY, "eval_batch: %d poses x %d pairs exceed the work-item index range; split the batch", npose, np);
    CU(ctx->d_items.reserve(want));
  }
  ep.items = ctx->d_items.p; ep.item_count = ctx->d_counters.p + 8; ep.item_cap = (int)std::min<size_t>(ctx->d_items.cap, 0x7fffffff);
  if (want_mask) CU(cudaMemsetAsync(d_mask, 0, (size_t)npose * mw * 8, ctx->stream));
  PROF_BEGIN(1); CU(launch_collide(ep, ctx->sm_count * 2, ctx->stream)); ctx->launches += 2; PROF_END(1);
  RecheckParams rp;
  memset(&rp, 0, sizeof rp);
  rp.ambig = ctx->d_ambig.p; rp.ambig_count = ctx->d_counters.p + 1; rp.ambig_cap = (int)ctx->d_ambig.cap;
  rp.nrb = nrb; rp.nb = nb; rp.fk = ctx->d_fk.p; rp.static_xf = P->d_static.p; rp.body_mesh = P->d_body_mesh.p;
  rp.meshes = ctx->d_meshes.p; rp.pairs = P->d_pairs.p; rp.legal = d_legal; rp.energy = d_energy; rp.pair_mask = d_mask;
  rp.mask_words = mw; rp.stats = ctx->d_stats.p;
  PROF_BEGIN(2); CU(launch_recheck(rp, ctx->stream)); ctx->launches++;
  if (want_energy) { CU(launch_compact(ep, ctx->stream)); ctx->launches++; }
  PROF_END(2);

  // K5 + K8: CONTACT_ENERGY over the compacted legal poses
  EnergyParams gp;
  memset(&gp, 0, sizeof gp);
  gp.npose = npose; gp.nrb = nrb; gp.nb = nb; gp.nvc = P->nvc; gp.nq = P->nhand; gp.object = P->object_e;
  gp.stage_nodes = ctx->stage_nodes;
  gp.stack_entries = energy_stack_entries_for(a->object >= 0 ? ctx->meshes[ctx->bodies[a->object].mesh].depth : 0);
  gp.fk = ctx->d_fk.p; gp.static_xf = P->d_static.p; gp.body_mesh = P->d_body_mesh.p; gp.meshes = ctx->d_meshes.p;
  gp.vcs = P->d_vcs.p; gp.legal = d_legal; gp.legal_list = ctx->d_legal_list.p; gp.legal_count = ctx->d_counters.p + 4;
  gp.energy = d_energy; gp.stats = ctx->d_stats.p; gp.work_counter = ctx->d_counters.p + 9;
  auto energy_blocks = [&](int contacts) {
    size_t smem = energy_block_smem_bytes(gp.stage_nodes, gp.stack_entries);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / smem));
    long long groups = ((long long)npose * contacts + 255) / 256;
    return (int)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * per_sm, groups));
  };
  if (want_energy) {
    CU(ctx->d_terms.reserve((size_t)npose * std::max(P->nvc, P->nhand) + 32));
    gp.terms = ctx->d_terms.p;
  }
  if (want_energy && !live) {
    if (P->nvc > 256) return fail(ctx, B2C_ECAPACITY, "eval_batch: more than 256 virtual contacts");
    gp.live = 0;
    PROF_BEGIN(4); CU(launch_energy(gp, energy_blocks(P->nvc), ctx->stream)); ctx->launches += 2; PROF_END(4);
  }

  // K4 (+ LIVE energy)
  float *d_dist = nullptr;
  const int nq = P->nhand;
  if (live || want_dist) {
    CU(ctx->d_dist.reserve((size_t)npose * nq));
    CU(ctx->d_pts.reserve((size_t)npose * nq * 6));
    // every query hands at least its winning pair to the fp64 refinement
    CU(ctx->d_near_dir.reserve((size_t)npose * nq + 1024));
    CU(ctx->d_near_ent.reserve(std::max<size_t>((size_t)1 << 21, 16 * (size_t)npose * nq)));     // lists of up to DNEAR_CAP pairs where fp32 ties
    d_dist = dev && want_dist ? a->link_dist : ctx->d_dist.p;
    DistParams dp;
    memset(&dp, 0, sizeof dp);
    dp.npose = npose; dp.nrb = nrb; dp.nb = nb; dp.nq = nq; dp.queries = P->d_queries.p; dp.fk = ctx->d_fk.p;
    dp.static_xf = P->d_static.p; dp.body_mesh = P->d_body_mesh.p; dp.meshes = ctx->d_meshes.p;
    dp.legal = want_dist ? nullptr : d_legal;       // LIVE alone only needs legal poses
    dp.dist = d_dist; dp.pts = ctx->d_pts.p; dp.work_counter = ctx->d_counters.p + 2;
    dp.ambig = ctx->d_ambig.p; dp.ambig_count = ctx->d_counters.p + 3; dp.ambig_cap = (int)ctx->d_ambig.cap; dp.stats = ctx->d_stats.p; dp.overflow = ctx->d_sticky.p;
    // near-ties and tiny distances are settled in fp64 (dist_refine_kernel): the closest POINTS feed CONTACT_LIVE
    dp.near_dir = ctx->d_near_dir.p; dp.near_ent = ctx->d_near_ent.p; dp.near_count = ctx->d_counters.p + 5;
    dp.near_cap_dir = (int)ctx->d_near_dir.cap; dp.near_cap_ent = (int)ctx->d_near_ent.cap;
    ctx->launches++;
    int dblocks = std::min(ctx->sm_count * 4, (npose * nq + 7) / 8);
    PROF_BEGIN(3); CU(launch_dist(dp, std::max(1, dblocks), ctx->stream)); ctx->launches++; PROF_END(3);
    RecheckParams rd;
    memset(&rd, 0, sizeof rd);
    rd.ambig = ctx->d_ambig.p; rd.ambig_count = ctx->d_counters.p + 3; rd.ambig_cap = (int)ctx->d_ambig.cap;
    rd.nrb = nrb; rd.nb = nb; rd.fk = ctx->d_fk.p; rd.static_xf = P->d_static.p; rd.body_mesh = P->d_body_mesh.p;
    rd.meshes = ctx->d_meshes.p; rd.pairs = P->d_queries.p; rd.dist = d_dist; rd.pts = ctx->d_pts.p; rd.nq = nq; rd.stats = ctx->d_stats.p;
    CU(launch_recheck(rd, ctx->stream)); ctx->launches++;
    if (live) {
      gp.live = 1; gp.pts = ctx->d_pts.p;
      PROF_BEGIN(4); CU(launch_energy(gp, energy_blocks(nq), ctx->stream)); ctx->launches += 2; PROF_END(4);
    }
  }

  // outputs
  if (!dev) {
    if (a->legal) CU(cudaMemcpyAsync(a->legal, d_legal, npose, cudaMemcpyDeviceToHost, ctx->stream));
    if (a->energy) CU(cudaMemcpyAsync(a->energy, d_energy, (size_t)npose * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_mask) CU(cudaMemcpyAsync(a->pair_mask, d_mask, (size_t)npose * mw * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_dist) CU(cudaMemcpyAsync(a->link_dist, d_dist, (size_t)npose * nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (want_tf) CU(cudaMemcpyAsync(a->body_tf, ctx->d_fkqt.p, (size_t)npose * nrb * 7 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaMemcpyAsync(ctx->last_stats, ctx->d_stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (!dev) {
    // host buffers: the call synchronises anyway, so a queue overflow is reported right here.  With device pointers the
    // call stays asynchronous; the kernels have answered fail-safe ("collides") and set the sticky word, which the next
    // synchronising call (b2c_check_overflow, b2c_anneal_read) turns into B2C_ECAPACITY.
    return b2c_check_overflow(ctx);
  }
  return B2C_OK;
}

int b2c_check_overflow(b2c_ctx *ctx) {
  if (!ctx) return B2C_EINVAL;
  cudaSetDevice(ctx->device);
  if (!ctx->d_sticky.p) { CU(cudaStreamSynchronize(ctx->stream)); return B2C_OK; }
  int word = 0;
  CU(cudaMemcpyAsync(&word, ctx->d_sticky.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (!word) return B2C_OK;
  CU(cudaMemsetAsync(ctx->d_sticky.p, 0, sizeof(int), ctx->stream));
  return fail(ctx, B2C_ECAPACITY, "queue overflow since the last check:%s%s%s%s -- the affected poses were reported as colliding; split the batch",
              (word & OVF_COLLIDE_QUEUE) ? " legality recheck queue" : "", (word & OVF_DIST_QUEUE) ? " distance recheck queue" : "",
              (word & OVF_ITEMS) ? " broad-phase work list" : "", (word & OVF_STACK) ? " traversal stack" : "");
}

} // extern "C"

// ---- contacts / region (K6, K7) ------------------------------------------------------------------
#include "contacts_api.inc"

// ---- batched annealing (A14) ---------------------------------------------------------------------
#include "anneal_api.inc"

// ---- batched autograsp (section 8(f).4) -------------------------------------------------------------
#include "autograsp_api.inc"
