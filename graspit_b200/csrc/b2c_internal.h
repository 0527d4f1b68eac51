// Internal data layouts shared by the host (api.cu, bvh.cpp) and the kernels (kernels.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "geom.cuh"
#include "bvh.h"

namespace b2c {

// triangle: three float4 (x,y,z,0), 48 B, 16-B aligned
struct Tri {
  float4 v[3];
};

struct MeshDev {
  const NodeQ *nodes;
  const Tri *tris;
  int nnodes;
  int ntri;
  float extent;      // |root centre|_1 + |root half|_1 : coordinate scale of the mesh
  int depth;
  float q0[3], qs[3]; // quantisation grid of the stored boxes (bvh.h)
  // closest-point seed grid (optional, built lazily for meshes used as the energy object): cell -> the triangle
  // closest to the cell's centre.  A query starts from the exact distance to that triangle instead of +inf.
  const int *seed;
  float g0[3], ginv[3];
  int gres[3];
};

// ---- forward-kinematics program ----------------------------------------
struct FkJoint {
  int dof;           // index into the pose's DOF vector (robot dof offset already added)
  int revolute;
  double ratio, theta, d;
  Qt tr4tr3;         // TRANSLATION(a,0,0) % AXIS_ANGLE_ROTATION(alpha, x)   (joint.cpp:68-88)
};
struct FkChain {
  Qt tran;
  int first_joint, njoints;
  int first_link, nlinks;
};
struct FkLink {
  int out_slot;      // slot in the per-pose output
  int last_joint;    // chain-local joint index after which the link pose is read
};
struct FkRobot {
  int base_slot, mount_slot;
  int parent;            // index into the program's robot list, -1 for the root
  int parent_last_slot;  // output slot of the parent chain's last link
  Qt parent_offset;
  int dof_ofs, ndof;
  int first_chain, nchains;
  Qt approach_inv;
  Qt approach;           // Robot::getApproachTran (SPACE_APPROACH needs it un-inverted too)
  int neg;
  int eg_ofs;            // offset into eg arrays: eg[neg*ndof], origin[ndof], norm[ndof], min[ndof], max[ndof]
};
struct FkProgram {
  const FkRobot *robots;
  const FkChain *chains;
  const FkJoint *joints;
  const FkLink *links;
  const double *egdata;
  int nrobots;
  int nslots;
  int ndof_total;
};

// ---- batched evaluation plan ---------------------------------------------------------------------
struct VcDev {
  int body;            // eval-body index
  int pad;
  double loc[3];
  double normal[3];
};

struct Ambig {         // triangle pair whose fp32 SAT was too close to call
  int pose, pair, tri_a, tri_b;
};

struct EvalParams {
  int npose;
  int nrb;             // robot bodies: transforms come from fk[pose][0..nrb)
  int nb;              // all eval bodies (robot bodies first, then static bodies)
  int npairs;
  int nvc;
  int object;          // eval-body index of the target object, -1 if none
  int mode_all;        // 1: evaluate every pair (ALL_COLLISIONS mask); 0: stop at the first hit
  int mask_words;      // 64-bit words per pose in pair_mask
  const Xf *fk;
  const Xf *static_xf;     // [nb - nrb]
  const int *body_mesh;    // [nb]
  const MeshDev *meshes;
  const int2 *pairs;       // [npairs] eval-body indices
  const VcDev *vcs;
  uint8_t *legal;
  float *energy;
  unsigned long long *pair_mask;
  const unsigned long long *pair_enable;   // optional [npose][mask_words]: pairs to test per pose (others are skipped)
  int *pose_counter;
  int *legal_list;         // poses without a collision, in completion order (for the energy kernel)
  int *legal_count;
  Ambig *ambig;
  int *ambig_count;
  int ambig_cap;
  unsigned long long *stats;   // [8]: box tests, tri-pair tests, point-box, point-tri, ambiguous, ...
  // two-phase legality (collide.cu): (pose, pair) work items that survived the root-box test
  int2 *items;
  int *item_count;
  int item_cap;
  int *overflow;           // sticky device word (b2c_ctx::d_sticky): OVF_* bits set when a queue drops an entry
};
enum { OVF_COLLIDE_QUEUE = 1, OVF_DIST_QUEUE = 2, OVF_ITEMS = 4, OVF_STACK = 8 };

// LIVE contacts / generic point-energy input: per pose, per contact: world location + world normal
struct LiveContact {
  double loc[3];
  double cn[3];
};

struct DistParams {
  int npose;
  int nrb, nb;
  int nq;                  // distance queries per pose
  const int2 *queries;     // [nq] (eval body a, eval body b)
  const Xf *fk;
  const Xf *static_xf;
  const int *body_mesh;
  const MeshDev *meshes;
  const uint8_t *legal;    // optional: skip poses with legal == 0
  const int *wlist;        // optional compact work list: entries pose * nq + query (autograsp bisection); length
  const int *wcount;       //   *wcount (device), capped at wcap; entries not listed keep their old dist value
  int wcap;
  float *dist;             // [npose][nq]  (-1 = interpenetration)
  double *pts;             // optional [npose][nq][6]: closest points, each in its own body frame
  int *work_counter;
  Ambig *ambig;
  int *ambig_count;
  int ambig_cap;
  // optional fp64 refinement of tiny distances: segment directory (query, offset, count) + triangle pairs
  int4 *near_dir;
  int2 *near_ent;
  int *near_count;         // [0] segments, [1] entries
  int near_cap_dir, near_cap_ent;
  unsigned long long *stats;
  int *overflow;           // sticky device word, see EvalParams
};

enum { STAT_BOX = 0, STAT_TRI = 1, STAT_PBOX = 2, STAT_PTRI = 3, STAT_AMBIG = 4, STAT_DBOX = 5, STAT_DTRI = 6, STAT_HITFIX = 7 };


struct FkParams {
  FkProgram prog;
  int npose;
  int input_mode;          // B2C_INPUT_* (include/b2c.h)
  double pos_params[3];    // SPACE_ELLIPSOID: a, b, c
  int stride;
  const float *states;
  const double *jstates;   // input_mode 2: per pose base (q wxyz, t) + one value per joint (program joint order), fp64
  Qt ref;                  // mRefTran for AA_EIGEN
  Xf *out;                 // [npose][nslots]
  double *out_qt;          // optional [npose][nslots][7]
};

struct RecheckParams {
  const Ambig *ambig;
  const int *ambig_count;
  int ambig_cap;
  int nrb, nb;
  const Xf *fk;
  const Xf *static_xf;
  const int *body_mesh;
  const MeshDev *meshes;
  const int2 *pairs;          // pair index -> (eval body a, eval body b)
  uint8_t *legal;
  float *energy;
  unsigned long long *pair_mask;
  int mask_words;
  float *dist;                // distance mode: dist[pose * nq + pair] = -1
  double *pts;
  int nq;
  unsigned long long *stats;
};

struct PointParams {
  int n;
  const double *points;     // [n][3] world
  const Xf *body_xf;        // [1]
  MeshDev mesh;
  double *out;              // [n][7]: dist, closest(3, world), normal(3) = unit(point -> closest)
};

struct EnergyParams {
  int npose;
  int nrb, nb;
  int nvc;                  // PRESET contacts per pose
  int nq;                   // LIVE contacts per pose (= hand bodies; query i = (hand body i, object))
  int object;
  int live;
  int stage_nodes;          // top-of-tree nodes of the object staged in shared memory by TMA (0 = off)
  int stack_entries;        // capacity of each warp's shared work stack (sized from the object tree's depth)
  const Xf *fk;
  const Xf *static_xf;
  const int *body_mesh;
  const MeshDev *meshes;
  const VcDev *vcs;
  const double *pts;        // LIVE: [npose][nq][6] closest points from the distance kernel
  const uint8_t *legal;
  const int *legal_list;    // optional compact list of legal poses (+ its device-side count)
  const int *legal_count;
  float *energy;
  float *terms;             // [legal poses x contacts] per-contact energy terms (scratch)
  int *work_counter;        // dynamic hand-out of 32-query groups (zeroed by the caller)
  unsigned long long *stats;
};

// ---- K6 / K7 ---------------------------------------------------------------------------------------
struct ContactRec {
  int query, tri_a, tri_b, feature;   // feature: 0-2 V(t2)-F(t1), 3-5 V(t1)-F(t2), 6-14 edge-edge
  double c[13];                       // b1_pos, b2_pos, b1_normal, b2_normal, distSq
};
struct ContactParams {
  int npose;
  int nrb, nb;
  int nq;
  double threshold;
  const int2 *queries;
  const Xf *fk;
  const Xf *static_xf;
  const int *body_mesh;
  const MeshDev *meshes;
  ContactRec *out;
  int out_cap;
  int *out_count;
  int *per_query;          // [npose * nq] contact candidates per query
  int *work_counter;
  int *overflow;           // sticky device word, see EvalParams
  int *work_list;          // optional [npose * nq]: filled by the launcher's cull pass with the queries whose root boxes come
  int *work_count;         //   within the threshold; the warps then take their work from this list
  int exhaustive;          // 1: no fp32 triangle-pair gate (every queued pair is enumerated in fp64) -- the check of the gate
};
struct RegionParams {
  MeshDev mesh;
  double point[3], normal[3], radius;
  unsigned char *flags;    // [ntri]
};
cudaError_t launch_contacts(const ContactParams &p, int blocks, cudaStream_t s);
cudaError_t launch_region(const RegionParams &p, cudaStream_t s);
struct RegionBatchParams {
  int nquery;
  const int *mesh_of;        // [nquery] mesh index of the query's body
  const MeshDev *meshes;
  const double *point;       // [nquery][3] body frame
  const double *normal;      // [nquery][3]
  const double *radius;      // [nquery]
  int2 *out;                 // (query, triangle) of every flagged triangle, arbitrary order
  int out_cap;
  int *out_count;
};
cudaError_t launch_region_batch(const RegionBatchParams &p, cudaStream_t s);
size_t region_lists_temp_bytes(int nhits);
cudaError_t launch_region_keep(const int2 *hits, int nhits, int n, const int *mesh_of, const MeshDev *meshes, int *first, int *keep,
                               int *pos, void *temp, size_t temp_bytes, cudaStream_t s);
cudaError_t launch_region_write(const int2 *hits, int nhits, int n, const int *mesh_of, const MeshDev *meshes, const double *points,
                                const int *first, const int *keep, const int *pos, double *out, long long cap, int *offsets, cudaStream_t s);
// K6b (contact_order.cu): per-candidate visit keys
struct OrderNode;
struct OrderParams {
  const ContactRec *recs;            // candidates grouped by query (local pose * np + pair)
  int n, np, nrb;
  const Xf *fk;                      // link transforms of the chunk's first pose onwards
  const Xf *static_xf;
  const int2 *pairs;
  const OrderNode *const *nodes;     // [engine body] reference-style OBB tree as the order rule reads it (obb_order.h)
  const int *const *tri_pos;         // [engine body] triangle -> depth-first leaf position
  unsigned long long *keys;          // [n][2] out: path bits, most significant first
  int *depth;                        // [n] out: path length (> 128: the key is truncated, the host replays the group)
  // rank pass (optional): recs / keys / depth / perm point at record `base` of the grouped array
  const int *offsets;                // [nqueries] first record of every query in the grouped array
  int nqueries, total, base;
  int *perm;                         // [n] out: perm[first of group - base + rank] = index inside the group; -1 = replay on the host
  // leaders (optional, needs offsets): candidates of one triangle pair share a path; only the first of them walks it
  int *leader;                       // [n] scratch: index of the candidate whose key this one copies (itself: a leader)
  int *leaders;                      // [n] scratch: dense list of the leaders
  int *nleaders;                     // [1] scratch
};
cudaError_t launch_contact_order(const OrderParams &p, cudaStream_t s);
// K6d: removeContactDuplicates on the device, per (pose, pair) group in visit order, + the survivors gathered into a compact
// array (contact_order.cu).  Queries [q0, q1) of the grouped array.
struct DedupeParams {
  const ContactRec *recs;            // grouped candidates (whole batch)
  const int *offsets;                // [nqueries] first record of every query
  int nqueries, total;
  int q0, q1;                        // this launch's queries
  int *perm;                         // [total] in: the group's visit order (index inside the group, -1 = unknown);
                                     //         out: the survivors' indices at the front of the group's segment
  int *nsurv;                        // [nqueries] out: survivors of the group; -1 = the host has to do this group
  int *soff;                         // [nqueries] scratch: exclusive scan of max(nsurv, 0) over [q0, q1)
  ContactRec *out;                   // survivors of [q0, q1), group after group, visit order inside a group
  int *out_total;                    // their number
  double thresh_sq;
  void *temp; size_t temp_bytes;
};
size_t contact_dedupe_temp_bytes(int nqueries);
cudaError_t launch_contact_dedupe(const DedupeParams &p, cudaStream_t s);
cudaError_t launch_contact_gather(const DedupeParams &p, int r0, int r1, cudaStream_t s);
size_t region_sort_temp_bytes(int n, int query_bits);
cudaError_t launch_region_sort(unsigned long long *pairs, unsigned long long *alt, int n, int query_bits, void *temp, size_t temp_bytes,
                               unsigned long long **sorted, cudaStream_t s);
size_t contact_group_temp_bytes(int nqueries);
cudaError_t launch_contact_group(const ContactRec *in, int n, const int *per_query, int nqueries, int *offsets, int *cursor,
                                 ContactRec *out, void *temp, size_t temp_bytes, cudaStream_t s);
cudaError_t launch_contact_digest(const int *per_query, const int *offsets, int nqueries, int np, int stride, int *out, int nslots, cudaStream_t s);

// ---- batched annealing (anneal.cu) ------------------------------------------------------------------
struct AnnealParams {
  int nchain, nvar;
  int step;                  // global step index (Philox counter)
  int chain_offset;          // global id of this context's first chain (multi-GPU: chains never migrate)
  unsigned long long seed;
  double YC, HC, NBR_ADJ, ERR_ADJ, T0;
  int YDIMS, HDIMS;
  const double *vmin, *vmax, *vjump;
  const unsigned char *circular;
  float *cur, *prop, *best;          // [nchain][nvar]
  float *cur_energy, *best_energy;   // [nchain]
  const float *energy;               // [nchain] energy of the proposals
  const unsigned char *legal;        // [nchain]
  int *k;                            // [nchain] per-chain annealing step (SimAnn::mCurrentStep)
  int *counters;                     // [0] jumps, [1] legal neighbours (totals)
  // incremental best list of this context (anneal_toplist_kernel)
  int *cand;                         // [nchain] chains whose best improved below the list threshold
  int *cand_count;
  float *top_thresh;                 // worst energy in the list (3e38 while it has free slots)
  int *top_chain;                    // [32] listed chains, -1 = empty
};
cudaError_t launch_anneal_toplist(const AnnealParams &p, int k, float *rows, cudaStream_t s);

// ---- best-row exchange between GPUs without a collective (anneal.cu) ---------------------------------------------------
// Every rank owns a ring of slots ring[depth][world][slot_floats] in its own HBM, mapped into its peers (CUDA IPC between
// processes, peer access inside one process).  A slot = rows x row_floats floats + a 16-byte tag field; tag = step + 1 of
// the rows it holds, negative while a writer is inside (sequence lock).
constexpr int EX_MAX_WORLD = 16;
struct ExchangeDev {
  float *ring[EX_MAX_WORLD];     // ring base of every rank, as seen from this device ([rank] = own ring)
  int world, rank, depth;
  int rows, row_floats, slot_floats;
};
cudaError_t launch_anneal_publish(const AnnealParams &p, const ExchangeDev &x, float *rows_local, cudaStream_t s);
cudaError_t launch_exchange_collect(const ExchangeDev &x, float *out, int *steps, cudaStream_t s);
cudaError_t launch_anneal_propose(const AnnealParams &p, cudaStream_t s);
cudaError_t launch_anneal_accept(const AnnealParams &p, cudaStream_t s);

constexpr int DSTACK_CAP = 448, DSTACK_PHYS = 512, DLEAF_CAP = 96;
#ifndef DNEAR_CAP_N
#define DNEAR_CAP_N 256   // near-tie list of a distance query; 64 overflowed once in 385 000 queries (one LIVE energy 1.08e-4 off), 128 and 256 never
#endif
constexpr int DNEAR_CAP = DNEAR_CAP_N, DCONF_CAP = 64;
constexpr int DIST_WARP_SMEM = DSTACK_PHYS * 8 + DLEAF_CAP * 8 + 3 * 96 + 96 + DNEAR_CAP * 8 + DCONF_CAP * 8 + DNEAR_CAP * 8 + 32;

// narrow phase of the legality test: per-warp shared memory = 32 fp64 relative transforms, 32 pair transforms,
// 32 item descriptors, the shared work stack and the leaf queue
#ifndef NSTACK_PHYS_N
#define NSTACK_PHYS_N 512
#endif
constexpr int NSTACK_PHYS = NSTACK_PHYS_N, NSTACK_CAP = NSTACK_PHYS_N - 64, NLEAF_CAP = 96;
constexpr int NARROW_WARP_SMEM = 32 * 96 + 32 * 96 + 32 * 16 + NSTACK_PHYS * 8 + NLEAF_CAP * 8 + 32 * 4;
cudaError_t launch_collide(const EvalParams &p, int narrow_blocks, cudaStream_t s);
cudaError_t launch_compact(const EvalParams &p, cudaStream_t s);

// cudaFuncSetAttribute applies to the CURRENT device only: remember per device the largest dynamic shared memory size
// already granted to a kernel (one process may hold contexts on several GPUs).
inline cudaError_t ensure_dyn_smem(const void *fn, size_t smem, size_t (&granted)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem <= granted[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) granted[dev] = smem;
  return e;
}

// launchers (kernels.cu)
cudaError_t launch_fk(const FkParams &p, cudaStream_t s);
cudaError_t launch_recheck(const RecheckParams &p, cudaStream_t s);
cudaError_t launch_point(const PointParams &p, cudaStream_t s);
cudaError_t launch_dist(const DistParams &p, int blocks, cudaStream_t s);
cudaError_t launch_energy(const EnergyParams &p, int blocks, cudaStream_t s);
cudaError_t launch_seed_grid(const MeshDev &m, int *seed, cudaStream_t s);
size_t energy_block_smem_bytes(int stage_nodes, int stack_entries);
int energy_stack_entries_for(int depth);

} // namespace b2c

// ---- batched autograsp (autograsp.cu): Robot::moveDOFToContacts for a batch of poses ----------------------------
namespace b2c {
constexpr int AG_MAXD = 32, AG_MAXJ = 32, AG_MAXL = 32, AG_MAXP = 512, AG_MW = AG_MAXP / 64;
enum { AG_CHECK = 1, AG_INTERP = 2, AG_DONE = 3 };
enum { AG_ST_MOVED = 1, AG_ST_INTERP_ERROR = 2, AG_ST_FAILSAFE = 4, AG_ST_STOPPED_AT_CONTACT = 8 };
struct AgDesc {
  int ndof, nj, nlinks, npairs;
  int mask_words;                  // 64-bit words of a pair mask: (npairs + 63) / 64
  int one_step;                    // desiredSteps == NULL: WorldElement::ONE_STEP for every DOF
  int stop_at_contact;
  int max_iters;                   // moveDOFToContacts failsafe (500)
  double threshold;                // Contact::THRESHOLD
  int dof_breakaway[AG_MAXD];      // 1: BreakAwayDOF, 0: RigidDOF
  double desired[AG_MAXD], step[AG_MAXD];
  int joint_dof[AG_MAXJ];
  double joint_ratio[AG_MAXJ];
  int link_last_joint[AG_MAXL];    // global joint index after which the link pose is read
  int link_first_joint[AG_MAXL];   // first joint of the link's chain
  short pair_link_a[AG_MAXP], pair_link_b[AG_MAXP];   // link index of each side of an active pair, -1 = not a link
};
// Per-pose state, stored field-major ([field][element][pose]) so that the threads of a warp touch consecutive
// addresses; AgState is the per-thread view of one pose.
template <typename T> struct AgCol {
  T *p;
  size_t stride;
  __host__ __device__ T &operator[](int i) const { return p[(size_t)i * stride]; }
};
struct AgStore {
  double *d;                   // (5 * AG_MAXD + 4 * AG_MAXJ + 2) x npose
  unsigned long long *u;       // 3 * AG_MW x npose
  unsigned char *b;            // 2 * AG_MAXJ x npose
  int *i;                      // 4 x npose
  int npose;
};
struct AgState {
  AgCol<double> q, desired, step, new_dof, initial_dof, jv, initial_jv, final_jv, break_val;
  double &t, &deltat;
  AgCol<unsigned long long> late, col, interest;
  AgCol<unsigned char> stopped, in_break;
  int &phase, &iters, &status, &ncols;
  __host__ __device__ AgState(const AgStore &s, int pose)
      : q{s.d + pose, (size_t)s.npose}, desired{s.d + (size_t)AG_MAXD * s.npose + pose, (size_t)s.npose},
        step{s.d + (size_t)2 * AG_MAXD * s.npose + pose, (size_t)s.npose},
        new_dof{s.d + (size_t)3 * AG_MAXD * s.npose + pose, (size_t)s.npose},
        initial_dof{s.d + (size_t)4 * AG_MAXD * s.npose + pose, (size_t)s.npose},
        jv{s.d + (size_t)(5 * AG_MAXD) * s.npose + pose, (size_t)s.npose},
        initial_jv{s.d + (size_t)(5 * AG_MAXD + AG_MAXJ) * s.npose + pose, (size_t)s.npose},
        final_jv{s.d + (size_t)(5 * AG_MAXD + 2 * AG_MAXJ) * s.npose + pose, (size_t)s.npose},
        break_val{s.d + (size_t)(5 * AG_MAXD + 3 * AG_MAXJ) * s.npose + pose, (size_t)s.npose},
        t(s.d[(size_t)(5 * AG_MAXD + 4 * AG_MAXJ) * s.npose + pose]),
        deltat(s.d[(size_t)(5 * AG_MAXD + 4 * AG_MAXJ + 1) * s.npose + pose]),
        late{s.u + pose, (size_t)s.npose}, col{s.u + (size_t)AG_MW * s.npose + pose, (size_t)s.npose},
        interest{s.u + (size_t)2 * AG_MW * s.npose + pose, (size_t)s.npose},
        stopped{s.b + pose, (size_t)s.npose}, in_break{s.b + (size_t)AG_MAXJ * s.npose + pose, (size_t)s.npose},
        phase(s.i[pose]), iters(s.i[(size_t)s.npose + pose]), status(s.i[(size_t)2 * s.npose + pose]),
        ncols(s.i[(size_t)3 * s.npose + pose]) {}
};
constexpr int AG_STORE_D = 5 * AG_MAXD + 4 * AG_MAXJ + 2, AG_STORE_U = 3 * AG_MW, AG_STORE_B = 2 * AG_MAXJ, AG_STORE_I = 4;
struct AgParams {
  int npose;
  int init;                        // 1: build the state from `states` and start the first step
  const AgDesc *desc;
  AgStore st;
  const float *states;             // init: raw states, base (q wxyz, t) + DOF values
  int stride;
  double *jstates;                 // [npose][7 + nj]: input of the joint-level FK
  const unsigned long long *pair_mask;   // [npose][mask_words] colliding pairs at the joint values of the last FK
  int mask_words;
  const float *dist;               // [npose][npairs] distances of the pairs listed in the previous round
  unsigned long long *pair_enable; // [npose][mask_words] pairs the next collision round has to test (0 once finished)
  int *wlist;                      // distance work list for the next round
  int *counters;                   // [0] poses not finished, [1] work-list length
  int wcap;
};
cudaError_t launch_autograsp_step(const AgParams &p, cudaStream_t s);
} // namespace b2c
