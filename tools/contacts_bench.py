#!/usr/bin/env python
"""BASELINE.json configs[2] as a throughput line: HumanHand20DOF vs flask, hands closed to contact by the batched
autograsp, then allContacts(0.2) for every pose in one b2c_contacts_batch call (+ per-contact friction-cone wrenches on
the host).  usage: python tools/contacts_bench.py [npose]   (GPU box)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspit_b200 import engine as E                                            # noqa: E402
from graspit_b200.engine import AG_INTERP_ERROR, B2C_INPUT_JOINTS, Engine, SceneBinding   # noqa: E402
from graspit_b200.formats.scenepack import load_pack                            # noqa: E402
from tests.common import aa_to_raw_states, object_body, sample_aa_states        # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    sc = load_pack("humanhand_flask")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    rid, obj = sb.robot_ids[0], sb.body_ids[object_body(sc)]
    r = sc.robots[0]
    rest = np.clip(r.dof_default, r.dof_min, r.dof_max)
    raw = np.zeros((0, 7 + r.n_dof), dtype=np.float32)
    seed = 100
    while len(raw) < n:
        pos, _ = sample_aa_states(sc, 4 * n, seed=seed, strata=("near",))
        cand = aa_to_raw_states(sc, pos, np.tile(rest[None, :], (len(pos), 1)).astype(np.float32))
        raw = np.concatenate([raw, cand[eng.eval_batch(rid, obj, cand)["legal"] == 1]])[:n]
        seed += 1
    t0 = time.perf_counter()
    ag = eng.auto_grasp(rid, r, raw)
    t_ag = time.perf_counter() - t0
    ok = (ag["status"] & AG_INTERP_ERROR) == 0
    js = np.concatenate([raw[ok, :7].astype(np.float64), ag["joints"][ok]], axis=1)
    eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS)             # warm-up at full size (staging buffers)
    t0 = time.perf_counter()
    res = eng.contacts_batch(rid, js, 0.2, input_mode=B2C_INPUT_JOINTS)
    t_c = time.perf_counter() - t0
    t0 = time.perf_counter()
    fixed = res["contacts"]
    wr = E.point_contact_wrenches(fixed[:20000], 1, 0.5, [0.0, 0.0, 0.0], 100.0) if len(fixed) else None
    t_w = (time.perf_counter() - t0) * (len(fixed) / max(1, min(len(fixed), 20000)))
    # neighbourhoods of both sides of every contact on an elastic body (findSoftNeighborhoods), one batched call
    pairs = eng.robot_pairs(rid)
    inv_b = {v: k for k, v in sb.body_ids.items()}
    y = np.array([[sc.bodies[inv_b[a]].youngs, sc.bodies[inv_b[b]].youngs] for a, b in pairs])
    soft = y.max(axis=1)[res["pair"]] > 0
    cs, pq = res["contacts"][soft], res["pair"][soft]
    pa = np.array(pairs)[pq]
    rad = np.array([E.soft_neighborhood_radius(*y[q]) for q in pq])
    qb = np.stack([pa[:, 0], pa[:, 1]], 1).reshape(-1)
    qp = np.stack([cs[:, 0:3], cs[:, 3:6]], 1).reshape(-1, 3)
    qn = np.stack([cs[:, 6:9], cs[:, 9:12]], 1).reshape(-1, 3)
    qr = np.repeat(rad, 2)
    eng.body_region_batch(qb, qp, qn, qr, flat=True)                         # warm-up at full size (pinned staging)
    t0 = time.perf_counter()
    flat, off = eng.body_region_batch(qb, qp, qn, qr, flat=True)
    t_r = time.perf_counter() - t0
    region = {"queries": int(len(qb)), "seconds": t_r, "queries_per_s": float(len(qb) / t_r) if len(qb) else 0.0,
              "mean_points": float(np.diff(off).mean()) if len(qb) else 0.0}
    # CPU baseline (reported, not a target): the reference's own allContacts on the same link poses, one thread, 64 poses
    cpu = None
    try:
        from tests.common import OracleScene                                     # test infrastructure (oracle/_ref)
        osc = OracleScene(sc, nthreads=1)
        bodies = eng.robot_bodies(rid)
        inv = {v: k for k, v in sb.body_ids.items()}
        tf = eng.eval_batch(rid, -1, js[:64], input_mode=B2C_INPUT_JOINTS, body_tf=True)["body_tf"]
        interest = [osc.ids[inv[b]] for b in bodies]
        t0 = time.perf_counter()
        for p in range(len(tf)):
            for k, b in enumerate(bodies):
                osc.w.set_transform(osc.ids[inv[b]], tf[p, k, :4], tf[p, k, 4:])
            osc.w.all_contacts(0.2, interest=interest)
        cpu = {"value": len(tf) / (time.perf_counter() - t0), "unit": "poses/s", "cores": 1, "kind": "reference",
               "sample": "64 of the same poses, GraspitCollision::allContacts compiled verbatim"}
    except Exception as e:                                                       # oracle not built on this box
        cpu = {"unavailable": str(e)[:80]}
    per_pose = np.diff(res["offsets"])
    print(json.dumps({"metric": "poses/s: allContacts(0.2) per pose, HumanHand20DOF vs flask (configs[2]), host buffers",
                      "npose": int(ok.sum()), "contacts_seconds": t_c, "value": float(ok.sum() / t_c),
                      "contacts_total": int(len(fixed)), "contacts_per_pose": float(per_pose.mean()),
                      "poses_with_contacts": float((per_pose > 0).mean()),
                      "autograsp_seconds": t_ag, "autograsps_per_s": float(len(raw) / t_ag), "autograsp_rounds": ag["rounds"],
                      "cone_wrenches_seconds_est": t_w, "active_pairs": len(eng.robot_pairs(rid)), "region_batch": region, "cpu_baseline": cpu}))


if __name__ == "__main__":
    main()
