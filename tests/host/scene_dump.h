// TEST -- the scene as plain numbers, written by tests/test_host_planner.py::_dump_scene, read by the C++ harnesses.
#ifndef B2C_TEST_SCENE_DUMP_H
#define B2C_TEST_SCENE_DUMP_H
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "b2c.h"

struct SceneDump {
  struct BodyD { double q[4], t[3]; std::vector<float> tri9; };
  std::vector<BodyD> bodies;
  std::vector<std::pair<int, int> > disabled;
  int base, mount, ndof;
  std::vector<double> dofMin, dofMax, dofVelocity;
  std::vector<int> dofBreakAway;
  std::vector<b2c_chain_desc> chains;
  std::vector<b2c_joint_desc> joints;
  std::vector<int> linkBody, linkLastJoint;       // scene body indices
  double approachQ[4], approachT[3];
  int neg;
  std::vector<double> eg, egOrigin, egNorm;
  struct Vc { int body; double loc[3], normal[3]; };
  std::vector<Vc> vcs;
  int object;
  double refQ[4], refT[3];
};

inline bool readSceneDump(const char *path, SceneDump *S)
{
  std::ifstream in(path);
  if (!in) return false;
  int nb; in >> nb;
  S->bodies.resize(nb);
  for (int i = 0; i < nb; i++) {
    SceneDump::BodyD &b = S->bodies[i];
    std::string file;
    in >> b.q[0] >> b.q[1] >> b.q[2] >> b.q[3] >> b.t[0] >> b.t[1] >> b.t[2] >> file;
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) return false;
    int n = 0; if (fread(&n, 4, 1, f) != 1) return false;
    b.tri9.resize(9 * (size_t)n);
    if (fread(b.tri9.data(), 4, b.tri9.size(), f) != b.tri9.size()) return false;
    fclose(f);
  }
  int nd; in >> nd;
  for (int i = 0; i < nd; i++) { int a, b; in >> a >> b; S->disabled.push_back(std::make_pair(a, b)); }
  in >> S->base >> S->mount >> S->ndof;
  S->dofMin.resize(S->ndof); S->dofMax.resize(S->ndof);
  for (int i = 0; i < S->ndof; i++) in >> S->dofMin[i];
  for (int i = 0; i < S->ndof; i++) in >> S->dofMax[i];
  int nchains; in >> nchains;
  S->chains.resize(nchains);
  for (auto &c : S->chains) in >> c.tran_q[0] >> c.tran_q[1] >> c.tran_q[2] >> c.tran_q[3] >> c.tran_t[0] >> c.tran_t[1] >> c.tran_t[2] >> c.first_joint >> c.njoints >> c.first_link >> c.nlinks;
  int nj; in >> nj;
  S->joints.resize(nj);
  for (auto &j : S->joints) in >> j.dof >> j.revolute >> j.ratio >> j.theta >> j.d >> j.a >> j.alpha;
  int nl; in >> nl;
  S->linkBody.resize(nl); S->linkLastJoint.resize(nl);
  for (auto &x : S->linkBody) in >> x;
  for (auto &x : S->linkLastJoint) in >> x;
  in >> S->approachQ[0] >> S->approachQ[1] >> S->approachQ[2] >> S->approachQ[3] >> S->approachT[0] >> S->approachT[1] >> S->approachT[2];
  in >> S->neg;
  S->eg.resize((size_t)S->neg * S->ndof); S->egOrigin.resize(S->ndof); S->egNorm.resize(S->ndof);
  for (auto &x : S->eg) in >> x;
  for (auto &x : S->egOrigin) in >> x;
  for (auto &x : S->egNorm) in >> x;
  int nvc; in >> nvc;
  S->vcs.resize(nvc);
  for (auto &v : S->vcs) in >> v.body >> v.loc[0] >> v.loc[1] >> v.loc[2] >> v.normal[0] >> v.normal[1] >> v.normal[2];
  in >> S->object;
  in >> S->refQ[0] >> S->refQ[1] >> S->refQ[2] >> S->refQ[3] >> S->refT[0] >> S->refT[1] >> S->refT[2];
  S->dofVelocity.resize(S->ndof); S->dofBreakAway.resize(S->ndof);
  for (int i = 0; i < S->ndof; i++) in >> S->dofVelocity[i];
  for (int i = 0; i < S->ndof; i++) in >> S->dofBreakAway[i];
  return (bool)in;
}

// the dumped world inside an engine context: body ids (scene index -> engine id), robot id
inline bool loadDumpIntoEngine(const SceneDump &S, int device, b2c_ctx **ctx, std::vector<int> *ids, int *robot)
{
  if (b2c_create(device, ctx) != B2C_OK) return false;
  ids->resize(S.bodies.size());
  for (size_t i = 0; i < S.bodies.size(); i++) {
    int mesh;
    if (b2c_mesh_create(*ctx, S.bodies[i].tri9.data(), (int)(S.bodies[i].tri9.size() / 9), &mesh) != B2C_OK ||
        b2c_body_create(*ctx, mesh, &(*ids)[i]) != B2C_OK) return false;
    b2c_body_set_transform(*ctx, (*ids)[i], S.bodies[i].q, S.bodies[i].t);
  }
  for (auto &p : S.disabled) b2c_pair_set_active(*ctx, (*ids)[p.first], (*ids)[p.second], 0);
  b2c_robot_desc d;
  memset(&d, 0, sizeof d);
  d.base_body = (*ids)[S.base]; d.mount_body = S.mount >= 0 ? (*ids)[S.mount] : -1; d.ndof = S.ndof;
  d.dof_min = S.dofMin.data(); d.dof_max = S.dofMax.data();
  d.nchains = (int)S.chains.size(); d.chains = S.chains.data();
  d.njoints = (int)S.joints.size(); d.joints = S.joints.data();
  d.nlinks = (int)S.linkBody.size();
  std::vector<int> lb(d.nlinks);
  for (int i = 0; i < d.nlinks; i++) lb[i] = (*ids)[S.linkBody[i]];
  d.link_body = lb.data(); d.link_last_joint = S.linkLastJoint.data();
  d.parent_robot = -1; d.parent_offset_q[0] = 1;
  for (int k = 0; k < 4; k++) d.approach_q[k] = S.approachQ[k];
  for (int k = 0; k < 3; k++) d.approach_t[k] = S.approachT[k];
  d.neg = S.neg; d.eg = S.eg.data(); d.eg_origin = S.egOrigin.data(); d.eg_norm = S.egNorm.data();
  std::vector<b2c_vcontact_desc> vcs(S.vcs.size());
  for (size_t i = 0; i < vcs.size(); i++) {
    vcs[i].body = (*ids)[S.vcs[i].body];
    for (int k = 0; k < 3; k++) { vcs[i].loc[k] = S.vcs[i].loc[k]; vcs[i].normal[k] = S.vcs[i].normal[k]; }
  }
  d.nvcontacts = (int)vcs.size(); d.vcontacts = vcs.data();
  return b2c_robot_define(*ctx, &d, robot) == B2C_OK;
}
#endif
