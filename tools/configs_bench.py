#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] as throughput lines on one GPU (host buffers, steady state):
  config 4: Staubli TX60L + Barrett vs floor / table / shelf, allCollisions(ALL) for 262 144 arm+hand configurations
  config 5: Barrett vs the mug subdivided to 220 800 triangles, 1 048 576 poses, legality + CONTACT_ENERGY (PRESET)
Each line carries a CPU sample of the reference's own backend on the same inputs (reported, not a target).
usage: python tools/configs_bench.py [4] [5]   (GPU box)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graspit_b200.engine import Engine, SceneBinding                            # noqa: E402
from graspit_b200.formats.scenepack import load_pack                            # noqa: E402
from tests.common import aa_to_raw_states, object_body, sample_aa_states        # noqa: E402


def _cpu(fn, n):
    try:
        t0 = time.perf_counter()
        fn()
        return {"value": n / (time.perf_counter() - t0), "unit": "poses/s", "cores": 1, "kind": "reference", "sample": f"{n} of the same inputs"}
    except Exception as e:
        return {"unavailable": str(e)[:80]}


def config4():
    sc = load_pack("staubli_barrett")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    arm, hand = sc.robots[0], sc.robots[1]
    rng = np.random.default_rng(1234)
    n = 262144
    base = np.tile(np.concatenate([arm.tran[0], arm.tran[1]])[None, :], (n, 1))
    dofs = np.concatenate([rng.uniform(arm.dof_min, arm.dof_max, size=(n, arm.n_dof)),
                           rng.uniform(hand.dof_min, hand.dof_max, size=(n, hand.n_dof))], 1)
    raw = np.concatenate([base, dofs], 1).astype(np.float32)
    rid = sb.robot_ids[0]
    eng.eval_batch(rid, -1, raw, pair_mask=True)
    t0 = time.perf_counter()
    res = eng.eval_batch(rid, -1, raw, pair_mask=True)
    dt = time.perf_counter() - t0

    def cpu():
        from tests.common import OracleScene
        OracleScene(sc, nthreads=1).pair_sets(0, raw[:128])
    print(json.dumps({"metric": "configurations/s: allCollisions(ALL, interest = robot), Staubli TX60L + Barrett vs obstacles (configs[3])",
                      "n": n, "seconds": dt, "value": n / dt, "collision_free": float(res["legal"].mean()),
                      "active_pairs": len(eng.robot_pairs(rid)), "triangles": int(sum(len(b.tris) for b in sc.bodies)),
                      "cpu_baseline": _cpu(cpu, 128)}))


def config5():
    sc = load_pack("barrett_mug_200k")
    eng = Engine(0)
    sb = SceneBinding(eng, sc)
    obj = object_body(sc)
    n = 1048576
    pos, dofs = sample_aa_states(sc, n, seed=1234)
    raw = aa_to_raw_states(sc, pos, dofs)
    rid, oid = sb.robot_ids[0], sb.body_ids[obj]
    eng.eval_batch(rid, oid, raw)
    t0 = time.perf_counter()
    res = eng.eval_batch(rid, oid, raw)
    dt = time.perf_counter() - t0

    def cpu():
        from tests.common import OracleScene
        OracleScene(sc, nthreads=1).eval(0, obj, raw[:2048], mode=0)
    print(json.dumps({"metric": "poses/s: legality + CONTACT_ENERGY (PRESET), Barrett vs 220 800-triangle mug (configs[4]), one GPU",
                      "n": n, "seconds": dt, "value": n / dt, "legal": float(res["legal"].mean()),
                      "object_triangles": int(len(sc.bodies[obj].tris)), "cpu_baseline": _cpu(cpu, 2048)}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["4", "5"]
    if "4" in which:
        config4()
    if "5" in which:
        config5()
