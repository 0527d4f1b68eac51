/*
 * b2c.h -- C ABI of the B200-native batched collision / proximity / contact engine.
 *
 * This is the drop-in boundary: every entry point replaces one
 * virtual of GraspIt!'s `class CollisionInterface`
 * (reference include/graspit/Collision/collisionInterface.h:44-185) or one step of the
 * `SearchEnergy::analyzeState` caller loop (reference src/EGPlanner/energy/searchEnergy.cpp:148-188),
 * and is what `B200Collision : public CollisionInterface` (graspit_b200/host/b200Collision.h,
 * INTEGRATION.md) binds inside GraspIt!.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - units: millimetres, radians; transforms are world-from-body, quaternion (w,x,y,z) + translation
 *   - every function returns 0 on success and a negative B2C_E* code on failure, never throws and
 *     never takes ownership of caller memory; b2c_last_error() describes the last failure
 *   - one context = one CUDA device + stream; calls on a context must be serialised by the caller
 *     (the reference keeps one CollisionInterface per World the same way)
 *   - there is NO CPU fallback: if no CUDA device is usable, b2c_create() fails
 */
#ifndef B2C_H
#define B2C_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2c_ctx b2c_ctx;

enum {
  B2C_OK = 0,
  B2C_EINVAL = -1,     /* bad argument / unknown id (reference: DBGA("GCOL: model not found")) */
  B2C_ECUDA = -2,      /* CUDA runtime failure (see b2c_last_error) */
  B2C_ECAPACITY = -3,  /* caller buffer or internal work queue too small */
  B2C_ENODEVICE = -4   /* no usable CUDA device: the engine has no CPU path */
};

/* ---- lifetime ----------------------------------------------------------------------------- */
int b2c_create(int device, b2c_ctx **out);
void b2c_destroy(b2c_ctx *ctx);
const char *b2c_last_error(const b2c_ctx *ctx);
/* CUDA stream the context launches on (cudaStream_t as void*), and the device ordinal */
void *b2c_stream(const b2c_ctx *ctx);
int b2c_device(const b2c_ctx *ctx);
/* number of kernel launches issued by this context so far (bench.py "gpu_launches") */
long long b2c_launch_count(const b2c_ctx *ctx);
/* launch on a caller-owned stream (cudaStream_t as void*; NULL restores the context's own stream), so
 * that callers who own device buffers (e.g. torch tensors) can order work and time it with their events */
int b2c_set_stream(b2c_ctx *ctx, void *stream);
/* per-kernel CUDA-event timing of the batched evaluation: when enabled, b2c_eval_batch brackets each kernel
 * with events on the launching stream; b2c_last_kernel_ms (which synchronises) returns the durations of the
 * last evaluation in ms: [0] fk, [1] eval (legality), [2] recheck, [3] distance (+ refine), [4] energy, [5..7] 0 */
int b2c_set_profiling(b2c_ctx *ctx, int enabled);
int b2c_last_kernel_ms(b2c_ctx *ctx, float ms[8]);

/* ---- geometry: replaces addBody / updateBodyGeometry / cloneBody / removeBody ---------------
 * collisionInterface.h:102-114, reference graspitCollision.cpp:200-305.  A mesh is an immutable
 * triangle soup + its BVH in device memory; a body is (mesh, transform, active flag).  Cloning a
 * body = creating a second body on the same mesh (reference CollisionModel::cloneModel,
 * collisionModel.cpp:466-474).  tri9: ntri x 9 floats (v1 v2 v3) -- the fp32 vertex values GraspIt!
 * obtains from Coin (reference src/body.cpp:1225-1249). */
int b2c_mesh_create(b2c_ctx *ctx, const float *tri9, int ntri, int *mesh_id);
int b2c_mesh_destroy(b2c_ctx *ctx, int mesh_id);
int b2c_body_create(b2c_ctx *ctx, int mesh_id, int *body_id);
int b2c_body_set_mesh(b2c_ctx *ctx, int body_id, int mesh_id);   /* updateBodyGeometry */
int b2c_body_remove(b2c_ctx *ctx, int body_id);

/* setBodyTransform (collisionInterface.h:116, graspitCollision.cpp:307-316) */
int b2c_body_set_transform(b2c_ctx *ctx, int body_id, const double q_wxyz[4], const double t[3]);
int b2c_body_get_transform(b2c_ctx *ctx, int body_id, double q_wxyz[4], double t[3]);

/* activateBody / activatePair / isActive (collisionInterface.h:123-131, graspitCollision.cpp:102-198).
 * b == -1 in b2c_pair_is_active asks about body a alone. */
int b2c_body_set_active(b2c_ctx *ctx, int body_id, int active);
int b2c_pair_set_active(b2c_ctx *ctx, int a, int b, int active);
int b2c_pair_is_active(b2c_ctx *ctx, int a, int b, int *active);

/* Thread rule of getActivePairs (graspitCollision.cpp:60-78; CollisionInterface::newThread / getThreadId,
 * collisionInterface.h:181-184): every body carries the id of the thread that added it (0 = master); a query
 * issued for thread T only pairs bodies of thread T or 0, and for T != 0 never two thread-0 bodies.
 * b2c_set_thread selects T for the queries that follow (default 0). */
int b2c_body_set_thread(b2c_ctx *ctx, int body_id, int thread_id);
int b2c_set_thread(b2c_ctx *ctx, int thread_id);

/* ---- single-pose queries at the bodies' current transforms -------------------------------- */

/* allCollisions (collisionInterface.h:141, graspitCollision.cpp:332-363).  interest == NULL means
 * all active pairs.  fast != 0 = FAST_COLLISION: *ncolliding is 0 or 1 and pairs_out is untouched.
 * Otherwise pairs_out receives (a,b) body ids, a < b, for up to cap colliding pairs. */
int b2c_all_collisions(b2c_ctx *ctx, int fast, const int *interest, int ninterest,
                       int *pairs_out, int cap, int *ncolliding);

/* bodyToBodyDistance (collisionInterface.h:161, graspitCollision.cpp:458-477): -1 when the bodies
 * interpenetrate.  p1/p2 reproduce the reference's output convention exactly, including its second
 * application of the inverse body transform; p1_local/p2_local (may be NULL)
 * receive the plain closest points in each body's own frame. */
int b2c_body_distance(b2c_ctx *ctx, int a, int b, double *dist, double p1[3], double p2[3],
                      double p1_local[3], double p2_local[3]);

/* pointToBodyDistance (collisionInterface.h:156, graspitCollision.cpp:435-456): world point in,
 * world closest point and unit vector point->closest out. */
int b2c_point_distance(b2c_ctx *ctx, int body, const double point[3], double *dist,
                       double closest[3], double normal[3]);

typedef struct b2c_contact {
  double b1_pos[3];     /* body-1 frame */
  double b2_pos[3];     /* body-2 frame */
  double b1_normal[3];  /* body-1 frame, points from body 2 into body 1 */
  double b2_normal[3];  /* body-2 frame */
  double distSq;        /* reference's mixed-frame value (collisionAlgorithms_inl.h:129-132) */
} b2c_contact;

/* contact (collisionInterface.h:149, graspitCollision.cpp:407-433) with the reference's duplicate
 * removal (3 mm) and perimeter compaction; raw != 0 returns the unprocessed candidate list. */
int b2c_contacts(b2c_ctx *ctx, int a, int b, double threshold, int raw,
                 b2c_contact *out, int cap, int *ncontacts);

/* allContacts (collisionInterface.h:146, graspitCollision.cpp:365-405): pair_ids gets (a,b) per
 * contacting pair, counts the contacts per pair, out the contacts of all pairs back to back. */
int b2c_all_contacts(b2c_ctx *ctx, double threshold, const int *interest, int ninterest,
                     int *pair_ids, int *counts, int cap_pairs, int *npairs,
                     b2c_contact *out, int cap_contacts, int *ncontacts);

/* bodyRegion (collisionInterface.h:171, graspitCollision.cpp:479-496): vertices (body frame) of all
 * triangles within radius of point whose normal does not oppose `normal`; the reference point itself
 * is appended last, as the reference does. */
int b2c_region(b2c_ctx *ctx, int body, const double point[3], const double normal[3], double radius,
               double *out_xyz, int cap, int *npoints);

/* bodyRegion for many (body, point, normal, radius) queries in one call -- the neighbourhoods of every contact of a
 * batch (findSoftNeighborhoods, reference contactSetting.cpp:114-139).  out_xyz receives the lists back to back, each as
 * b2c_region returns it (vertices of the flagged triangles in triangle order, repeats removed, the query point last);
 * offsets[n + 1] gives the first point of each list.  *npoints is the total; lists beyond cap points are truncated. */
int b2c_region_batch(b2c_ctx *ctx, int n, const int *bodies, const double *points, const double *normals,
                     const double *radii, double *out_xyz, int cap, int *offsets, int *npoints);

typedef struct b2c_obb {
  double t[3];
  double q_wxyz[4];
  double half[3];
} b2c_obb;

/* getBoundingVolumes (collisionInterface.h:175, graspitCollision.cpp:499-509): the reference-style
 * oriented-box hierarchy (PCA fit, mean-vertex split) at `depth`, for display and for the reference's
 * own unit test; independent of the device BVH used for queries. */
int b2c_bounding_volumes(b2c_ctx *ctx, int body, int depth, b2c_obb *out, int cap, int *nboxes);

/* Context-free host utilities (no CUDA): the reference-style box hierarchy of a triangle soup, and the
 * reference's contact-set post-processing (removeContactDuplicates at `duplicate_threshold`, then
 * compactContactSet; collisionInterface.cpp:74-287) applied in place; *n is updated. */
int b2c_obb_hierarchy(const float *tri9, int ntri, int depth, b2c_obb *out, int cap, int *nboxes);
int b2c_contacts_postprocess(b2c_contact *contacts, int *n, double duplicate_threshold, int compact);

/* Order in which the reference's ContactCallback reaches a set of triangle pairs (ta[i] of model 1, tb[i] of model 2):
 * its dual-tree recursion over the two OBB hierarchies splits the node with the larger box volume and descends first
 * into the child with the smaller bboxDistanceSq (reference src/Collision/Graspit/collisionAlgorithms.cpp:42-130,
 * include/graspit/bbox_inl.h:269-413).  removeContactDuplicates and compactContactSet depend on this order
 * (src/Collision/collisionInterface.cpp:74-287), so b2c_contacts / b2c_all_contacts / b2c_contacts_batch put the
 * device's candidates into it before post-processing; this context-free entry point exposes the replay for tests.
 * q*, t*: world transforms of the two models.  order[n] receives indices into (ta, tb); entries that share a triangle
 * pair keep their input order. */
int b2c_contact_visit_order(const float *tri9_a, int ntri_a, const float *tri9_b, int ntri_b, const double qa_wxyz[4],
                            const double ta_xyz[3], const double qb_wxyz[4], const double tb_xyz[3], const int *ta,
                            const int *tb, int n, int *order);

/* Contact frames and point-contact friction cones, the last step of World::findAllContacts -> addContacts
 * (reference src/contactSetting.cpp:193-221).  Context-free host utilities (closed form, a few hundred flops per contact).
 * side = 1: the contact as seen from body 1 (b1_pos, b1_normal), side = 2: its mate on body 2.
 *   b2c_check_contact_normals  checkContactNormals (contactSetting.cpp:54-81): flips normals that point the wrong way
 *                              given the bodies' world transforms; *nflipped = contacts changed
 *   b2c_contact_frames         Contact::Contact (src/contact/contact.cpp:67-99): frame7 = [n][q wxyz, t], x/y = tangents
 *   b2c_point_contact_wrenches PointContact (PCWF) friction cone: 8 edges of the unit circle
 *                              (pointContact.cpp:44-56, contact.cpp:180-216) turned into boundary wrenches
 *                              f = n + cof (e0 tx + e1 ty), tau = (r x f + n cof e5) / max_radius, r = pos - cog
 *                              (contact.cpp:124-165); wrench48 = [n][8][force 3, torque 3] */
int b2c_check_contact_normals(b2c_contact *contacts, int n, const double q1_wxyz[4], const double t1[3],
                              const double q2_wxyz[4], const double t2[3], int *nflipped);
int b2c_contact_frames(const b2c_contact *contacts, int n, int side, double *frame7);
int b2c_point_contact_wrenches(const b2c_contact *contacts, int n, int side, double cof, const double cog[3],
                               double max_radius, double *wrench48);

/* Soft-finger contacts (SFCL), the elastic branch of addContacts (reference src/contactSetting.cpp:114-221,
 * src/contact/softContact.cpp:22-125,186-358, include/graspit/FitParabola.h).  Context-free host utilities; the
 * neighbourhoods they consume are the output of b2c_region (bodyRegion) around each contact.
 * Neighbourhood lists are ragged: xyz = points back to back (3 doubles each, body frame), offsets[n + 1] = first point
 * of each contact's list (offsets[0] = 0).
 *   b2c_soft_neighborhood_radius  findSoftNeighborhoods' radius (contactSetting.cpp:129-133), mm
 *   b2c_soft_contact_merge        mergeSoftNeighborhoods (contactSetting.cpp:141-184): contacts whose body-1
 *                                 neighbourhoods share a vertex are one contact; the closer one (world distance between
 *                                 its two points, contactSetting.cpp:39-47) survives and takes the union of both lists.
 *                                 In place: contacts, *n, both xyz arrays and both offset arrays are rewritten.
 *   b2c_soft_contact_fit          SoftContact::SoftContact + FitPoints (softContact.cpp:22-58,186-199): paraboloid
 *                                 z = a x^2 + b y^2 + c xy through the neighbourhood in the contact frame, then
 *                                 RotateParaboloid -> radii of curvature r1, r2 (mm, -1 = flat) and the rotation angle.
 *                                 As in the reference, the neighbourhood is ROTATED into the contact frame but not
 *                                 re-centred on the contact (softContact.cpp:41-43).
 *   b2c_soft_contact_wrenches     setUpFrictionEdges of one side (softContact.cpp:69-117): CalcRelPhi, CalcRprimes,
 *                                 CalcContact_Mattress(5 N, 3 mm mattress, youngs = max of the two bodies, Pa) -> contact
 *                                 ellipse -> friction ellipsoid with latitudes {1,5,8,5,1} (20 edges,
 *                                 contact.cpp:180-216) -> boundary wrenches (contact.cpp:124-165).
 *                                 wrench120 = [n][20][force 3, torque 3]; ellipse6 (optional) = [n][major axis m,
 *                                 minor axis m, relPhi, r1', r2' (mm), max torque / (mu N) m].  q_this / q_mate: world
 *                                 rotations of the body on `side` and of the other body. */
typedef struct b2c_soft_fit {
  double a, b, c;
  double r1, r2;
  double rot_angle;
} b2c_soft_fit;
double b2c_soft_neighborhood_radius(double youngs1, double youngs2);
int b2c_soft_contact_merge(b2c_contact *contacts, int *n, const double q1_wxyz[4], const double t1[3],
                           const double q2_wxyz[4], const double t2[3], double *nghbd1_xyz, int *offsets1,
                           double *nghbd2_xyz, int *offsets2, int *nmerged);
int b2c_soft_contact_fit(const b2c_contact *contacts, int n, int side, const double *nghbd_xyz, const int *offsets,
                         b2c_soft_fit *fit);
int b2c_soft_contact_wrenches(const b2c_contact *contacts, int n, int side, const b2c_soft_fit *fit_this,
                              const b2c_soft_fit *fit_mate, const double q_this_wxyz[4], const double q_mate_wxyz[4],
                              double youngs, double cof, const double cog[3], double max_radius, double *wrench120,
                              double *ellipse6);

/* ---- batched pose evaluation: FK -> legal -> proximity -> CONTACT_ENERGY --------------------
 * Replaces, for whole batches of candidate states, the loop
 *   HandObjectState::execute -> Robot::setTran / forceDOFVals -> KinematicChain::fwdKinematics
 *   SearchEnergy::legal -> World::noCollision(hand) -> allCollisions(FAST, NULL, &handBodies)
 *   ContactEnergy::energy -> [World::findVirtualContacts] -> VirtualContact::getObjectDistanceAndNormal
 * (reference searchState.cpp:515-527, kinematicChain.cpp:543-560, searchEnergy.cpp:104-188,
 *  contactEnergy.cpp:9-55, world.cpp:1478-1508,1618-1651). */

typedef struct b2c_joint_desc {
  int dof;            /* index into the robot's DOF vector */
  int revolute;       /* 1 revolute, 0 prismatic */
  double ratio;       /* coupling ratio: joint = ratio * dof */
  double theta;       /* DH theta at joint value 0 (revolute: the offset c) */
  double d;           /* DH d at joint value 0 (prismatic: the offset c) */
  double a;
  double alpha;
} b2c_joint_desc;

typedef struct b2c_chain_desc {
  double tran_q[4], tran_t[3];   /* chain base w.r.t. robot base */
  int first_joint, njoints;      /* range in the joints array */
  int first_link, nlinks;        /* range in the link arrays */
} b2c_chain_desc;

typedef struct b2c_vcontact_desc {
  int body;                      /* body id the contact is attached to */
  double loc[3];                 /* body frame */
  double normal[3];              /* body frame */
} b2c_vcontact_desc;

typedef struct b2c_robot_desc {
  int base_body;                 /* body id of the palm / base link */
  int mount_body;                /* body id of the mount piece or -1 */
  int ndof;
  const double *dof_min, *dof_max;         /* [ndof] */
  int nchains;
  const b2c_chain_desc *chains;
  int njoints;
  const b2c_joint_desc *joints;
  int nlinks;
  const int *link_body;          /* [nlinks] body id per link, chain-major */
  const int *link_last_joint;    /* [nlinks] joint (chain-local) after which the link pose is read */
  int parent_robot;              /* robot id this one is attached to, or -1 */
  int parent_chain;
  double parent_offset_q[4], parent_offset_t[3];   /* child base w.r.t. parent chain end */
  /* search-state mapping (B2C_INPUT_AA_EIGEN) */
  double approach_q[4], approach_t[3];     /* Robot::getApproachTran */
  int neg;                                 /* number of eigengrasps */
  const double *eg;                        /* [neg x ndof] normalised eigengrasp directions */
  const double *eg_origin, *eg_norm;       /* [ndof] */
  int nvcontacts;
  const b2c_vcontact_desc *vcontacts;      /* CONTACT_PRESET virtual contacts */
} b2c_robot_desc;

int b2c_robot_define(b2c_ctx *ctx, const b2c_robot_desc *desc, int *robot_id);

enum {
  B2C_INPUT_RAW = 0,       /* per pose: base (qw qx qy qz tx ty tz) + ndof DOF values (root robot first,
                              then attached robots' DOFs in robot id order); stride = 7 + total dofs */
  B2C_INPUT_AA_EIGEN = 1,  /* per pose: Tx Ty Tz theta phi alpha + neg eigengrasp amplitudes
                              (SPACE_AXIS_ANGLE + POSE_EIGEN, searchStateImpl.cpp:83-95,156-168) */
  B2C_INPUT_JOINTS = 2,    /* per pose: base (qw qx qy qz tx ty tz) + one value per JOINT (chain-major, the order of
                              b2c_robot_desc.joints), as DOUBLES: `states` points to doubles, stride counts doubles.
                              For postures whose joints no longer equal ratio x DOF (break-away fingers after
                              b2c_autograsp_batch: feed joint_out) */
  /* the other search-state layouts of HandObjectState (reference src/EGPlanner/searchStateImpl.cpp): position variables first,
   * posture variables after them; hand transform = mRefTran % getCoreTran() (searchState.cpp:515-527) */
  B2C_INPUT_ELLIPSOID_EIGEN = 3,  /* SPACE_ELLIPSOID (:218-252): beta gamma tau dist, then neg eigengrasp amplitudes;
                                     ellipsoid parameters a b c in b2c_eval_args.position_params */
  B2C_INPUT_APPROACH_EIGEN = 4,   /* SPACE_APPROACH (:266-274): dist, wrist 1, wrist 2, then neg eigengrasp amplitudes */
  B2C_INPUT_COMPLETE_DOF = 5      /* SPACE_COMPLETE + POSE_DOF (:44-49,125-135): Tx Ty Tz Qw Qx Qy Qz, then the root robot's DOFs */
};

enum {
  B2C_EVAL_PRESET = 0,       /* CONTACT_PRESET: file-defined virtual contacts */
  B2C_EVAL_LIVE = 1,         /* CONTACT_LIVE: one virtual contact per link from bodyToBodyDistance */
  B2C_EVAL_PAIRMASK = 2,     /* also evaluate ALL_COLLISIONS and return the colliding-pair bitmask */
  B2C_EVAL_LINKDIST = 4,     /* also return bodyToBodyDistance(link, object) per hand body */
  B2C_EVAL_DEVICE_PTRS = 8   /* states/outputs are device pointers (already resident in HBM) */
};

typedef struct b2c_eval_args {
  int robot;                 /* root robot (hand, or arm with the hand attached) */
  int object;                /* body id of the target object (energy / distances); -1 = legality only */
  int npose;
  int input_mode;
  const float *states;       /* npose x stride floats (B2C_INPUT_JOINTS: doubles behind the same pointer) */
  int stride;
  double ref_q[4], ref_t[3]; /* B2C_INPUT_AA_EIGEN: mRefTran (normally the object pose) */
  int flags;
  uint8_t *legal;            /* [npose] 1 = no collision among pairs touching the robot */
  float *energy;             /* [npose] CONTACT_ENERGY; 0 when illegal */
  uint64_t *pair_mask;       /* [npose x mask_words] bit i = i-th active pair collides (B2C_EVAL_PAIRMASK) */
  float *link_dist;          /* [npose x (nlinks + 1)] (B2C_EVAL_LINKDIST): finger links chain-major, then the palm -- the
                              * first nlinks + 1 entries of b2c_robot_bodies (the mount piece carries no contact) */
  double *body_tf;           /* optional [npose x nbodies x 7]: FK result (q wxyz, t), host or device */
  double position_params[3]; /* B2C_INPUT_ELLIPSOID_EIGEN: a, b, c (all zero = the reference's 80, 80, 160) */
} b2c_eval_args;

int b2c_eval_batch(b2c_ctx *ctx, const b2c_eval_args *args);

/* bodies moved by a robot (and robots attached to it) in the order used by link_dist / body_tf:
 * finger links chain-major, then the base, then the mount piece (World::findVirtualContacts order,
 * reference world.cpp:1623-1639), children appended after their parent. */
/* Batched calls with B2C_EVAL_DEVICE_PTRS (and the annealer built on them) stay asynchronous, so they cannot return a
 * queue overflow themselves.  The kernels answer fail-safe -- a triangle pair the fp32 filter could not decide and that
 * no longer fits the fp64 recheck queue counts as a COLLISION -- and set a sticky word on the device; this call
 * synchronises the context's stream, returns B2C_ECAPACITY (and clears the word) if any queue overflowed since the last
 * check, B2C_OK otherwise.  Host-buffer b2c_eval_batch, b2c_anneal_read and b2c_autograsp_batch call it themselves. */
int b2c_check_overflow(b2c_ctx *ctx);

int b2c_robot_bodies(b2c_ctx *ctx, int robot, int *bodies, int cap, int *nbodies);
/* the active pairs (a,b) that a batched evaluation of `robot` tests, in pair_mask bit order */
int b2c_robot_pairs(b2c_ctx *ctx, int robot, int *pairs, int cap, int *npairs);

/* ---- batched contacts: allContacts(threshold, interest = the robot's bodies) for every pose of a batch -----------
 * Per pose what World::findContacts / findAllContacts ask of the backend for the hand (reference world.cpp:1682-1760,
 * graspitCollision.cpp:365-405): every active pair that touches a robot body, contact() with the reference's duplicate
 * removal (3 mm) and perimeter compaction per pair (raw != 0: the unprocessed candidates).  Contacts come back grouped
 * by pose, then by pair; contact_pair indexes b2c_robot_pairs(robot).  *ncontacts is the total found; when it exceeds
 * cap_contacts the arrays hold the first cap_contacts. */
typedef struct b2c_contacts_batch_args {
  int robot;
  int npose;
  int input_mode;            /* B2C_INPUT_* */
  const void *states;        /* floats (RAW, AA_EIGEN) or doubles (JOINTS) */
  int stride;
  double ref_q[4], ref_t[3]; /* B2C_INPUT_AA_EIGEN */
  double threshold;          /* Contact::THRESHOLD = 0.2 mm in the reference */
  int raw;
  b2c_contact *contacts;     /* [cap_contacts] */
  int *contact_pose;         /* [cap_contacts] */
  int *contact_pair;         /* [cap_contacts] */
  int cap_contacts;
  int *ncontacts;
  int *pose_offsets;         /* optional [npose + 1]: first contact of each pose in the (untruncated) listing */
} b2c_contacts_batch_args;
int b2c_contacts_batch(b2c_ctx *ctx, const b2c_contacts_batch_args *args);

/* ---- batched autograsp: Robot::moveDOFToContacts for whole batches of hand poses -------------------------------
 * Replaces, per pose, Hand::autoGrasp -> Robot::moveDOFToContacts -> jumpDOFToContact -> getJointValuesFromDOF /
 * interpolateJoints / stopJointsFromLink and the distance filter of World::findContacts (reference src/robot.cpp:
 * 2427-2462, 1708-1815, 1593-1706, 1507-1560, 1418-1472, 1474-1489; src/world.cpp:1682-1707; RigidDOF / BreakAwayDOF in
 * src/dof.cpp:497-514, 540-648).  Every pose starts from its own base pose and DOF values (a RAW row, as in
 * b2c_eval_batch; the DOFs are taken as freshly forced: no break-away memory) and is closed towards `desired` in
 * increments of `steps` until every DOF reached its target, is stopped by contacts, or the reference's 500-step
 * failsafe hits.  Collisions are found among the active pairs that touch a still-moving link (allCollisions with the
 * reference's interest list); a colliding step is bisected on bodyToBodyDistance of the colliding pairs exactly as
 * interpolateJoints does (accept when 0 < distance < contact_threshold / 2).
 * The start posture must be free of contacts on the links (the state HandObjectState::execute leaves:
 * Body::setTran breaks a moved link's contacts, body.cpp:928-931); then Link::contactsPreventMotion inside
 * getJointValuesFromDOF cannot fire and is not evaluated.  Robots with other robots attached are refused.
 * Hand::autoGrasp(speedFactor): desired[d] = speedFactor * defaultVelocity[d] >= 0 ? max : min,
 * steps[d] = defaultVelocity[d] * speedFactor * 0.01 (Robot::AUTO_GRASP_TIME_STEP). */
enum {
  B2C_AG_MOVED = 1,                /* moveDOFToContacts' return value */
  B2C_AG_INTERP_ERROR = 2,         /* "Robot joint interpolation error": the start was already colliding; DOFs not updated */
  B2C_AG_FAILSAFE = 4,             /* more than 500 steps */
  B2C_AG_STOPPED_AT_CONTACT = 8
};
typedef struct b2c_autograsp_args {
  int robot;
  int npose;
  const float *states;             /* npose x stride floats: base (qw qx qy qz tx ty tz) + ndof DOF values */
  int stride;
  const double *desired;           /* [ndof] desiredVals */
  const double *steps;             /* [ndof] desiredSteps (0: the DOF does not move); NULL = WorldElement::ONE_STEP */
  const unsigned char *dof_breakaway;  /* [ndof] DOF type: 0 = RigidDOF (<dof type="r">), 1 = BreakAwayDOF ("b"), 2 = CompliantDOF ("c":
                                       * every joint moves on its own, dof.cpp:830-856; the joints' `ratio` in b2c_robot_desc must
                                       * then hold CompliantDOF::getStaticRatio, dof.cpp:779-795) */
  int stop_at_contact;
  double contact_threshold;        /* Contact::THRESHOLD; <= 0 selects the reference's 0.2 mm */
  int max_rounds;                  /* safety bound on host-loop rounds; <= 0 selects the default */
  double *dof_out;                 /* [npose x ndof] final DOF values */
  double *joint_out;               /* optional [npose x njoints] final joint values (chain-major joint order) */
  int *status;                     /* optional [npose x 2]: B2C_AG_* flags, moveDOFToContacts iterations */
  unsigned char *stopped_out;      /* optional [npose x njoints] stoppedJoints bits (1 positive, 2 negative direction) */
  int *rounds;                     /* optional: host-loop rounds used (each = state step + FK + collide + distances) */
} b2c_autograsp_args;
int b2c_autograsp_batch(b2c_ctx *ctx, const b2c_autograsp_args *args);

/* ---- batched simulated annealing: K independent chains on the device ------------------------
 * Replaces, for K chains at once, SimAnnPlanner::mainLoop -> SimAnn::iterate (reference
 * src/EGPlanner/simAnnPlanner.cpp:111-148, simAnn.cpp:96-261): per step every chain draws one neighbour of its
 * current state (Ingber distribution at the neighbour temperature), the batch is evaluated by b2c_eval_batch
 * (B2C_INPUT_AA_EIGEN states), and the Metropolis rule accepts or rejects per chain.  A chain whose neighbour is
 * illegal keeps its state and its step counter (the reference's FAIL).  Random numbers are Philox4x32-10 keyed by
 * `seed` with counter (chain_offset + chain, step, draw), so trajectories are independent of batch layout and of
 * how chains are split over GPUs. */
typedef struct b2c_anneal_desc {
  int robot, object;
  int nchain;
  int nvar;                         /* 6 + number of eigengrasps */
  const double *vmin, *vmax, *vjump;/* [nvar] SearchVariable mMinVal, mMaxVal, mMaxJump (searchStateImpl.cpp:50-80,147-155) */
  const uint8_t *circular;          /* [nvar] SearchVariable::isCircular */
  double ref_q[4], ref_t[3];        /* mRefTran */
  double YC, HC, NBR_ADJ, ERR_ADJ, T0;   /* SimAnnParams (simAnnParams.cpp) */
  int YDIMS, HDIMS, K0;
  unsigned long long seed;
  int chain_offset;                 /* global index of this context's first chain */
  int flags;                        /* B2C_EVAL_PRESET or B2C_EVAL_LIVE */
} b2c_anneal_desc;

/* initial_states: [nchain][nvar] fp32 (host).  Initial energy is 1e8 and the step counter K0, as
 * SimAnnPlanner::resetParameters sets them (simAnnPlanner.cpp:63-70). */
int b2c_anneal_create(b2c_ctx *ctx, const b2c_anneal_desc *desc, const float *initial_states, int *anneal_id);
int b2c_anneal_destroy(b2c_ctx *ctx, int anneal_id);
/* nsteps annealing steps, asynchronous on the context's stream */
int b2c_anneal_step(b2c_ctx *ctx, int anneal_id, int nsteps);
/* host copies (any pointer may be NULL): current and best-so-far state / energy per chain, per-chain step counters;
 * totals = {accepted jumps, legal neighbours, steps taken, neighbours evaluated}.  Synchronises. */
int b2c_anneal_read(b2c_ctx *ctx, int anneal_id, float *cur, float *cur_energy, float *best, float *best_energy,
                    int *k, long long totals[4]);
/* device pointers of the per-chain best / current rows, for callers that reduce or exchange them on the device
 * (the multi-GPU best-list all-gather) */
int b2c_anneal_device_ptrs(b2c_ctx *ctx, int anneal_id, void **best, void **best_energy, void **cur, void **cur_energy);
/* this context's k (<= 32) best chains by best-so-far energy as rows [k][nvar + 1] = (state || energy) in DEVICE
 * memory (empty slots: energy 3e38), maintained incrementally from the chains that improved since the last call;
 * asynchronous on the context's stream.  It is the payload of the per-step multi-GPU all-gather that feeds the
 * planner's best list (SimAnnPlanner::mainLoop, simAnnPlanner.cpp:126-143).  Always call it with the same k. */
int b2c_anneal_best_rows(b2c_ctx *ctx, int anneal_id, int k, void *rows_device);
/* test hooks: the neighbours the next step would draw (no state change), and overwriting chain state */
/* Multi-GPU: exchange of every GPU's BEST_LIST_SIZE best (state || energy) rows for the planner's best-list merge
 * (reference src/EGPlanner/egPlanner.cpp:527-557) WITHOUT a collective.  Every rank owns a small ring of
 * slots in its own HBM which its peers map (CUDA IPC between processes, peer access inside one process); after a step
 * b2c_anneal_publish brings the context's best list up to date and stores its rows straight into every peer's ring over
 * NVLink under a sequence lock -- one kernel, no NCCL launch, no rendezvous, a rank never waits for another one.
 * b2c_exchange_collect copies out, per source rank, the newest complete rows that have landed (steps[src] = annealing
 * steps that source had completed + 1, 0 = nothing yet).  Rows only ever improve, so a reader that is a step behind or ahead merges
 * the same list a little earlier or later.
 *   create:  one exchange per context; ipc_handle_out (64 bytes, optional) is what the other PROCESSES need,
 *            ring_out (optional) what other contexts of THIS process need
 *   connect: ipc_handles = world x 64 bytes (all ranks, own entry ignored), or peer_rings = world device pointers
 *            (+ peer_devices to enable peer access) when all contexts live in one process
 *   publish: rows_device (optional) also receives this rank's rows [rows x (nvar + 1)]
 *   collect: rows = world x rows x (nvar + 1) floats on the host (pinned for async != 0: the copy is then only enqueued) */
int b2c_exchange_create(b2c_ctx *ctx, int rank, int world, int depth, int rows, int row_floats, void *ipc_handle_out, void **ring_out);
int b2c_exchange_connect(b2c_ctx *ctx, const void *ipc_handles, void *const *peer_rings, const int *peer_devices);
int b2c_anneal_publish(b2c_ctx *ctx, int anneal_id, void *rows_device);
int b2c_exchange_collect(b2c_ctx *ctx, float *rows, int *steps, int async);
int b2c_exchange_destroy(b2c_ctx *ctx);

int b2c_anneal_propose(b2c_ctx *ctx, int anneal_id, float *proposals);
int b2c_anneal_write(b2c_ctx *ctx, int anneal_id, const float *cur, const float *cur_energy, const int *k, int step);

/* per-context counters of the last batched evaluation: box tests, triangle-pair tests,
 * point-triangle tests, fp64 rechecks (explanatory L2-level traffic) */
int b2c_last_stats(b2c_ctx *ctx, unsigned long long stats[8]);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H */
